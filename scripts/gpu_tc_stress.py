"""Stress the tensor-core lifting kernel: many launches over varied view shapes (segment structures), results compared
with the FP32 kernel each time.  A protocol bug shows up as a trap (launch failure) or a mismatch."""
import os, sys, time
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
torch.manual_seed(0)
shapes = [(16, 256, 768), (1, 8, 52), (3, 17, 53), (2, 100, 104), (5, 64, 200), (48, 32, 96), (7, 9, 300), (1, 1024, 64), (2, 300, 51)]
t0 = time.time(); worst = 0.0; n = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    for shp in shapes:
        src = torch.rand(*shp, device=dev) - 0.5
        din = torch.rand(*shp, device=dev) - 0.5
        o1, o2 = torch.empty_like(src), torch.empty_like(src)
        for k in range(4):
            ops.lift_step([(src, din, o1)], blobs[k % len(blobs)], 1.0 if k % 2 == 0 else -1.0, 0.1, False, "tc")
            ops.lift_step([(src, din, o2)], blobs[k % len(blobs)], 1.0 if k % 2 == 0 else -1.0, 0.1, False, "fp32")
            worst = max(worst, (o1 - o2).abs().max().item()); n += 1
src = torch.rand(16, 256, 768, device=dev) - 0.5; din = torch.rand_like(src); o1 = torch.empty_like(src)
for _ in range(2000):
    ops.lift_step([(src, din, o1)], blobs[0], 1.0, 0.1, False)
torch.cuda.synchronize()
print(f"{n} compared launches + 2000 back-to-back launches in {time.time() - t0:.1f} s, worst |tc - fp32| = {worst:.2e}")
assert worst < 5e-6
