import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model as om
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops, _lib
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import LiftingBasedNeuralWaveletv4
dev = "cuda:0"
cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=4)
torch.manual_seed(1337)
net = LiftingBasedNeuralWaveletv4(cfg).to(dev).eval()
blobs = net.waveletForward[0]._blobs()
lib = _lib.load()
ops.set_lift_mode("tc")
src = torch.rand(16, 256, 768, device=dev) - 0.5
din = torch.rand(16, 256, 768, device=dev) - 0.5
out = torch.empty_like(src)
buf = torch.zeros(17 * 8, dtype=torch.int64, device=dev)
for _ in range(2): ops.lift_step([(src, din, out)], blobs[0], 1.0, 0.1, False)
lib.ll_lift_set_debug_buffer(buf.data_ptr())
ops.lift_step([(src, din, out)], blobs[0], 1.0, 0.1, False)
torch.cuda.synchronize()
lib.ll_lift_set_debug_buffer(None)
b = buf.cpu().view(17, 8)
t0 = int(b[b > 0].min())
names = {"E": ["din-ld", "mma-wait", "E-A", "bar1", "E-B", "conv4", "end", "after-barrier"],
         "S": ["loads", "conv1", "conv4", "skip", "-", "-", "end", "after-barrier"],
         "M": ["start", "free-wait", "issued", "-", "-", "-", "end", "after-barrier"]}
for w in range(17):
    g = "E" if w < 8 else ("S" if w < 16 else "M")
    row = [(names[g][k], int(b[w, k]) - t0) for k in range(8) if int(b[w, k]) > 0]
    print(f"warp {w:2d} {g}: " + "  ".join(f"{n}={v}" for n, v in row))
