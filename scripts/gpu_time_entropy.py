"""Time the conditioned2ZT entropy model at BASELINE config-3 subband shapes (one colour plane)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model as om
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
    DWTConditioned2EntropyLayerZTsepSubbands
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = "cuda:0"
res = {}
for prec in ("bf16", "fp32"):
    cfg = om.default_cfg(dwtlevels=4, ctx_precision=prec)
    torch.manual_seed(1337)
    em = DWTConditioned2EntropyLayerZTsepSubbands(cfg).to(dev).eval()
    torch.manual_seed(0)
    xe = torch.randn(B, 1, 32, 48, device=dev) * 4
    xo = [torch.randn(B, 3, 256 >> l, 384 >> l, device=dev) * 4 for l in range(4)]
    n = 3 if prec == "bf16" else 1
    with torch.no_grad():
        em(xe, xo)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        e0.record()
        for _ in range(n):
            out = em(xe, xo)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    bits = float(out[0].double().sum() + sum(s.double().sum() for s in out[1]))
    res[prec] = {"ms_per_plane": ms, "launches": (ops.launch_count() - l0) // n, "MP_per_s_3planes": B * 512 * 768 / 1e6 / (3 * ms * 1e-3), "bits": bits}
    print(prec, json.dumps(res[prec]))
print("bpp rel diff bf16 vs fp32:", abs(res["bf16"]["bits"] - res["fp32"]["bits"]) / res["fp32"]["bits"])
json.dump(res, open("gpurun_out/entropy_timing.json", "w"), indent=1)
