"""Last conv of the Berk scaling network (ll_nhwc_split_conv3) timed alone: 96 -> 3 on 8 / 32 planes of 256x384, 32 -> 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
torch.manual_seed(0)
dev = "cuda:0"
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B, C, Co, H, W in ((8, 96, 3, 256, 384), (32, 96, 3, 256, 384), (32, 96, 3, 128, 192), (32, 32, 1, 32, 48)):
    z = torch.randn(B, H, W, 2 * C, device=dev)
    w = torch.randn(Co, C, 3, 3, device=dev) * 0.05
    b = torch.randn(Co, device=dev)
    ms = t(lambda: ops.nhwc_split_conv3(z, w, b))
    print(f"tail conv {C}->{Co} on {B}x{H}x{W}: {ms:.3f} ms, {z.numel() * 4 / ms / 1e6:.0f} GB/s read")
