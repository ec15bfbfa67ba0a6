"""One eval forward of conditioned2ZTsepSubbands (tensor-core context path) on 16 planes of 512x768 subband shapes: the
command the layer's ncu launch list / captures are taken from."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import config as C
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
    DWTConditioned2EntropyLayerZTsepSubbands
dev = "cuda:0"
cfg = C.default_config(dwtlevels=4)
torch.manual_seed(1337)
em = DWTConditioned2EntropyLayerZTsepSubbands(cfg).to(dev).eval()
torch.manual_seed(0)
B = 16
xe = torch.randn(B, 1, 32, 48, device=dev) * 4
xo = [torch.randn(B, 3, 256 >> l, 384 >> l, device=dev) * 4 for l in range(4)]
with torch.no_grad():
    for _ in range(2):
        em(xe, xo)
torch.cuda.synchronize(); print("ok")
