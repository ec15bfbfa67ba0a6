"""lift_step_tc_kernel: role ablation and per-warp step timeline through the probe build (csrc/probe/lift_dbg.cu)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib, ops
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import config as C
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import LiftingBasedNeuralWaveletv4
P = ctypes.c_void_p
probe = _lib.load_probe()
real = _lib.load()
dev = "cuda:0"
cfg = C.default_config(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=4)
torch.manual_seed(1337)
net = LiftingBasedNeuralWaveletv4(cfg).to(dev).eval()
blobs = net.waveletForward[0]._blobs()
src = torch.rand(16, 256, 768, device=dev) - 0.5
din = torch.rand(16, 256, 768, device=dev) - 0.5
out = torch.empty_like(src)
# route ops.lift_step's library call through the probe build
probe.ll_lift_step.restype = ctypes.c_int
probe.ll_lift_step.argtypes = real.ll_lift_step.argtypes
class Shim:
    def __getattr__(self, k):
        return getattr(probe if k == "ll_lift_step" else real, k)
_lib._lib = Shim()
MODE = sys.argv[1] if len(sys.argv) > 1 else "tc"
def run():
    ops.lift_step([(src, din, out)], blobs[0], 1.0, 0.1, False, MODE)
def timed(n=10):
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ONLY = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else None   # a subset of the switch sets (own process per set)
for name, bits in (("full", 0), ("no MMA", 1), ("no E-B", 2), ("no conv1", 4), ("no conv4", 8), ("no E-A", 16),
                   ("no workers' math", 2 | 4 | 8 | 16), ("nothing", 31)):
    if ONLY is not None and bits not in ONLY:
        continue
    probe.ll_dbg_lift_switches(bits)
    print(f"{name:20s} {timed():.3f} ms", flush=True)
probe.ll_dbg_lift_switches(0)
buf = torch.zeros(20, 8, dtype=torch.int64, device=dev)
probe.ll_dbg_lift_stamp_buffer(ops.ptr(buf))
run(); torch.cuda.synchronize()
probe.ll_dbg_lift_stamp_buffer(None)
t = buf.cpu()
if int(t[19][4]) and int(t[19][3]):
    print(f"SM clock under load: {(int(t[19][5]) - int(t[19][0])) / (int(t[19][4]) - int(t[19][3])):.3f} GHz; "
          f"{(int(t[19][5]) - int(t[19][0])) / 200:.0f} cycles per step (steps 40..240 of CTA 0)")
    t[19][3:6] = 0
t0 = int(t[t > 0].min())
names = {0: "epi0", 1: "epi1", 4: "epi4(conv3 drain)", 6: "epi6", 8: "conv1", 12: "conv4", 19: "MMA"}
for w, nm in names.items():
    print(f"warp {w:2d} {nm:18s}", [int(v) - t0 if v else None for v in t[w]])
