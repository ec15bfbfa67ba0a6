"""One eval forward of the ZTBlock entropy layer at config-3 subband shapes (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model as om
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
    DWTConditioned2EntropyLayerZTBlock
dev = "cuda:0"; B, H, W, L = 16, 512, 768, 4
cfg = om.default_cfg(entropy_layer="DWTConditioned2EntropyLayerZTBlock", dwtlevels=L)
torch.manual_seed(1337); em = DWTConditioned2EntropyLayerZTBlock(cfg).to(dev).eval()
xe = torch.randn(B, 1, H >> L, W >> L, device=dev) * 4
xo = [torch.randn(B, 3, H >> (l + 1), W >> (l + 1), device=dev) * 4 for l in range(L)]
with torch.no_grad():
    for _ in range(2):
        em(xe, xo)
torch.cuda.synchronize(); print("ok")
