"""Timing experiments of the TC lifting kernel: disable phases via the debug bits."""
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
src = torch.rand(16, 256, 768, device=dev) - 0.5
din = torch.rand(16, 256, 768, device=dev) - 0.5
out = torch.empty_like(src)
def t(mode):
    _lib.check(lib.ll_lift_set_mode(mode))
    for _ in range(2): ops.lift_step([(src, din, out)], blobs[0], 1.0, 0.1, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.lift_step([(src, din, out)], blobs[0], 1.0, 0.1, False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
px = 16 * 256 * 768
for name, mode in [("fp32 SIMT", 0), ("tc full", 1), ("tc no MMA", 1 | (1 << 8)), ("tc no E-B", 1 | (2 << 8)), ("tc no conv1", 1 | (4 << 8)),
                   ("tc no conv4", 1 | (8 << 8)), ("tc no E-A", 1 | (16 << 8)), ("tc no workers' math (E-A,E-B,conv1,conv4 off)", 1 | (30 << 8)),
                   ("tc MMA only off + all math off", 1 | (31 << 8))]:
    ms = t(mode)
    print(f"{name:50s} {ms:8.3f} ms   {ms * 1e-3 * 1.9e9 * 148 / px:7.1f} SM-cycles/px")
_lib.check(lib.ll_lift_set_mode(1))
