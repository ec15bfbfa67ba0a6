import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib
dev = "cuda:0"; lib = _lib.load(); N, h, w = 48, 512, 768
x = torch.rand(N, h, w, device=dev); ll = torch.empty(N, h // 2, w // 2, device=dev); yh = torch.empty(N, 3, h // 2, w // 2, device=dev); xr = torch.empty_like(x)
sp = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    lib.ll_dwt97_fwd_level(x.data_ptr(), h * w, ll.data_ptr(), h * w // 4, yh.data_ptr(), 3 * h * w // 4, N, h, w, sp)
    lib.ll_dwt97_inv_level(ll.data_ptr(), h * w // 4, yh.data_ptr(), 3 * h * w // 4, xr.data_ptr(), h * w, N, h, w, sp)
torch.cuda.synchronize(); print("ok")
