import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import SubbandAutoEncoderBerk
torch.manual_seed(0)
ae = SubbandAutoEncoderBerk(3).to("cuda:0").eval()
x = torch.randn(8, 3, 256, 384, device="cuda:0")
with torch.no_grad():
    for _ in range(2):
        y = ae.encode(x)
torch.cuda.synchronize(); print("ok")
