import os, sys, time
sys.path.insert(0, '/root/repo')
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
mode = int(sys.argv[1])
shape = tuple(int(v) for v in sys.argv[2].split('x'))
lib.ll_lift_set_mode(ops.LIFT_TC | (mode << 8))
src = torch.rand(*shape, device=dev) - 0.5
din = torch.rand(*shape, device=dev) - 0.5
out = torch.empty_like(src)
torch.cuda.synchronize()
t = time.time()
try:
    ops.lift_step([(src, din, out)], blobs[0], 1.0, 0.1, False)
    torch.cuda.synchronize()
    print("mode", mode, shape, "ok", time.time() - t)
except Exception as e:
    print("mode", mode, shape, "FAIL after", time.time() - t, str(e)[:80])
