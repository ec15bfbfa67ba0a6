import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model as om
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import LiftingBasedNeuralWaveletv4
dev = "cuda:0"
cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=4)
torch.manual_seed(1337)
net = LiftingBasedNeuralWaveletv4(cfg).to(dev).eval()
blobs = net.waveletForward[0]._blobs()
src = torch.rand(16, 256, 768, device=dev) - 0.5
din = torch.rand(16, 256, 768, device=dev) - 0.5
out = torch.empty_like(src)
for _ in range(3): ops.lift_step([(src, din, out)], blobs[0], 1.0, 0.1, False)
torch.cuda.synchronize(); print("ok")
