import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
torch.manual_seed(0)
dev = "cuda:0"
N = 96
x = torch.randn(8, 3, 256, 384, device=dev)
w0 = torch.randn(N, 3, 3, 3, device=dev) * 0.2
gp = ops.pack_tf32_weight((torch.rand(N, N, device=dev) * 0.01 + 0.1 * torch.eye(N, device=dev)).reshape(N, N, 1, 1).contiguous())
b, beta = torch.zeros(N, device=dev), torch.ones(N, device=dev)
for _ in range(3):
    z = ops.conv3_gdn_head(x, w0, b, gp, beta)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    z = ops.conv3_gdn_head(x, w0, b, gp, beta)
e1.record(); torch.cuda.synchronize()
print("head kernel ms", e0.elapsed_time(e1) / 10)
