"""Time ll_igemm_conv on plc-shaped problems (run on the GPU box)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops

def run(B, H, W, cin, cout, taps, out_mode):
    dev = "cuda:0"
    k = 3 if taps == 9 else 1
    kpad = (cin + 63) // 64 * 64
    x = torch.zeros(B, H, W, kpad, dtype=torch.bfloat16, device=dev)
    x[..., :cin] = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, device=dev) * 0.02
    b = torch.randn(cout, device=dev)
    wp = ops.pack_igemm_weight(w)
    out = torch.empty(B, cout, H, W, device=dev) if out_mode == "nchw" else None
    onh = torch.empty(B, H, W, (cout + 63) // 64 * 64, dtype=torch.bfloat16, device=dev) if out_mode == "nhwc" else None
    f = lambda: ops.igemm_conv(x, wp, b, cout, out=out, out_nhwc=onh)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * B * H * W * cin * cout * taps
    npad = (cout + 15) // 16 * 16
    flops_pad = 2.0 * B * H * W * kpad * npad * taps
    return {"B": B, "H": H, "W": W, "cin": cin, "cout": cout, "taps": taps, "out": out_mode, "ms": ms,
            "tflops_useful": flops / ms / 1e9, "tflops_issued": flops_pad / ms / 1e9}

res = []
for cfg in [(8, 256, 384, 243, 243, 9, "nchw"), (8, 256, 384, 243, 243, 9, "nhwc"), (8, 128, 192, 243, 243, 9, "nchw"),
            (8, 256, 384, 162, 162, 1, "nhwc"), (8, 256, 384, 256, 256, 9, "nhwc")]:
    r = run(*cfg); res.append(r); print(json.dumps(r))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/igemm_timing.json", "w"), indent=1)
