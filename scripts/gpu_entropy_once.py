"""One eval pass of the conditioned2ZT entropy model (bf16 context path) at config-3 subband shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model as om
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
    DWTConditioned2EntropyLayerZTsepSubbands
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = "cuda:0"
cfg = om.default_cfg(dwtlevels=4)
torch.manual_seed(1337)
em = DWTConditioned2EntropyLayerZTsepSubbands(cfg).to(dev).eval()
xe = torch.randn(B, 1, 32, 48, device=dev) * 4
xo = [torch.randn(B, 3, 256 >> l, 384 >> l, device=dev) * 4 for l in range(4)]
with torch.no_grad():
    for _ in range(reps):
        em(xe, xo)
torch.cuda.synchronize()
print("ok")
