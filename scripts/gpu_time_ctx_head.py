"""Small-Cin context convs (plc head, masked csc) at the level-0 chunk shape: fp32 SIMT kernels vs im2col + 1-tap igemm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
B, H, W = 16, 256, 384
con = torch.round(torch.randn(B, 3, H // 2, W // 2, device=dev) * 30)
q = torch.round(torch.randn(B, 3, H, W, device=dev) * 20)
wh = (torch.rand(243, 3, 3, 3, device=dev) * 2 - 1) * 0.3
bh = torch.rand(243, device=dev) - 0.5
wc = (torch.rand(243, 1, 5, 5, device=dev) * 2 - 1) * 0.2
mask = torch.ones(5, 5, device=dev); mask[2, 2:] = 0; mask[3:] = 0
wc = wc * mask
bc = torch.rand(243, device=dev) - 0.5
whp = torch.zeros(243, 128, 1, 1, device=dev); whp[:, :81, 0, 0] = ops.split_bf16_weight(wh.reshape(243, 27))
pk_head = ops.pack_igemm_weight(whp, npad=256, kpad=128)
wcp = []
for g in range(3):
    wg = torch.zeros(81, 64, 1, 1, device=dev)
    wg[:, :36, 0, 0] = ops.split_bf16_weight(wc[81 * g:81 * (g + 1), 0].reshape(81, 25)[:, :12])
    wcp.append(ops.pack_igemm_weight(wg, npad=128, kpad=64))
pk_csc = torch.stack(wcp).contiguous()
g_in = torch.empty(B, H, W, 640, dtype=torch.bfloat16, device=dev)
t = torch.empty(B, H, W, 256, dtype=torch.bfloat16, device=dev)
a = ops.ctx_im2col(con, q)
def timed(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print(f"SIMT head     {timed(lambda: ops.ctx_conv_nhwc(con, wh, bh, upsample2=True, lrelu=True, region=256)):.3f} ms")
print(f"SIMT csc      {timed(lambda: ops.ctx_conv_nhwc(q, wc, bc, groups=3, live_taps=12, out=g_in, coff=256, co_group=81, co_gstride=128, region=384)):.3f} ms")
print(f"im2col        {timed(lambda: ops.ctx_im2col(con, q)):.3f} ms")
print(f"igemm head    {timed(lambda: ops.igemm_conv(a, pk_head, bh, 243, lrelu=True, out_nhwc=t, koff=[[0, 64]])):.3f} ms")
print(f"igemm csc     {timed(lambda: ops.igemm_conv(a, pk_csc, bc, 81, out_nhwc=g_in, nhwc_coff=256, nhwc_gstride=128, koff=[[128], [192], [256]])):.3f} ms")
