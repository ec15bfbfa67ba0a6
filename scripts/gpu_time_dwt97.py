import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops, _lib
dev = "cuda:0"; lib = _lib.load()
N = 48
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
sp = torch.cuda.current_stream().cuda_stream
for (h, w) in [(512, 768), (256, 384), (128, 192), (64, 96)]:
    x = torch.rand(N, h, w, device=dev)
    ll = torch.empty(N, h // 2, w // 2, device=dev); yh = torch.empty(N, 3, h // 2, w // 2, device=dev)
    xr = torch.empty_like(x)
    f = lambda: lib.ll_dwt97_fwd_level(x.data_ptr(), h * w, ll.data_ptr(), h * w // 4, yh.data_ptr(), 3 * h * w // 4, N, h, w, sp)
    g = lambda: lib.ll_dwt97_inv_level(ll.data_ptr(), h * w // 4, yh.data_ptr(), 3 * h * w // 4, xr.data_ptr(), h * w, N, h, w, sp)
    mf, mi = t(f), t(g)
    by = 8.0 * N * h * w
    print(f"level {h}x{w}: fwd {mf*1e3:7.1f} us {by/mf/1e6:7.0f} GB/s | inv {mi*1e3:7.1f} us {by/mi/1e6:7.0f} GB/s | PR {(xr-x).abs().max().item():.2e}")
# config-3 sized batch (64 images x 3 planes) at level 0
N = 192; h, w = 512, 768
x = torch.rand(N, h, w, device=dev)
ll = torch.empty(N, h // 2, w // 2, device=dev); yh = torch.empty(N, 3, h // 2, w // 2, device=dev)
xr = torch.empty_like(x)
mf, mi = t(f), t(g)
by = 8.0 * N * h * w
print(f"batch-192 level {h}x{w}: fwd {mf*1e3:7.1f} us {by/mf/1e6:7.0f} GB/s | inv {mi*1e3:7.1f} us {by/mi/1e6:7.0f} GB/s | PR {(xr-x).abs().max().item():.2e}")
mc = t(lambda: xr.copy_(x))
print(f"torch copy 302MB: {mc*1e3:.1f} us {2*x.numel()*4/mc/1e6:.0f} GB/s")
# plain device copy for reference
a = torch.rand(N, 512, 768, device=dev); b = torch.empty_like(a)
mc = t(lambda: b.copy_(a))
print(f"torch copy 75.5MB: {mc*1e3:.1f} us {2*a.numel()*4/mc/1e6:.0f} GB/s")
a = torch.rand(512 * 1024 * 1024 // 4, device=dev); b = torch.empty_like(a)
mc = t(lambda: b.copy_(a), 10)
print(f"torch copy 512MB: {mc*1e3:.1f} us {2*a.numel()*4/mc/1e6:.0f} GB/s")
