"""A few launches of the plc-shaped tcgen05 implicit GEMM (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
dev = "cuda:0"
B, H, W = 8, 256, 384
x = torch.zeros(B, H, W, 256, dtype=torch.bfloat16, device=dev)
x[..., :243] = torch.randn(B, H, W, 243, device=dev).to(torch.bfloat16)
w = torch.randn(243, 243, 3, 3, device=dev) * 0.02
b = torch.randn(243, device=dev)
wp = ops.pack_igemm_weight(w, npad=256, kpad=256)
out = torch.empty(B, H, W, 256, dtype=torch.bfloat16, device=dev)
for _ in range(4):
    ops.igemm_conv(x, wp, b, 243, out_nhwc=out)
torch.cuda.synchronize()
print("ok")
