import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
torch.manual_seed(0)
z = torch.randn(32, 256, 384, 192, device="cuda:0")
w = torch.randn(3, 96, 3, 3, device="cuda:0") * 0.05
b = torch.randn(3, device="cuda:0")
for _ in range(2):
    ops.nhwc_split_conv3(z, w, b)
torch.cuda.synchronize()
print("ok")
