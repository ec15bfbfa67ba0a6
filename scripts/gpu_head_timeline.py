"""Per-tile timeline of the head instance of igemm_tf32_gdn_pair_kernel (probe build with LL_TIMELINE): cycles between the
hand-offs of CTA 0's MMA thread and of its first epilogue warp."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib, ops
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.ops import ptr, stream_ptr
lib = _lib.load_probe()
P = ctypes.c_void_p
lib.ll_conv3_gdn_head.restype = ctypes.c_int
lib.ll_conv3_gdn_head.argtypes = [P] * 5 + [ctypes.c_int] * 6 + [P, P]
def timed(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
torch.manual_seed(0)
N, iC = 96, 3
x = torch.randn(8, iC, 256, 384, device="cuda:0")
w0 = torch.randn(N, iC, 3, 3, device="cuda:0") * 0.2
gp = ops.pack_tf32_weight((torch.rand(N, N, device="cuda:0") * 0.01 + 0.1 * torch.eye(N, device="cuda:0")).reshape(N, N, 1, 1).contiguous())
b = torch.zeros(N, device="cuda:0"); beta = torch.ones(N, device="cuda:0")
sz = torch.empty(8, 256, 384, 2 * N, device="cuda:0")
run = lambda: _lib.check_probe(lib.ll_conv3_gdn_head(ptr(x), ptr(w0), ptr(b), ptr(gp), ptr(beta), 8, iC, 256, 384, N, 0, ptr(sz), stream_ptr()))
for ns in (1, 0):
    lib.ll_probe_set_nostore(ns)
    print(f"head conv {iC}->{N} + GDN, E2 global stores {'off' if ns else 'on'}: {timed(run):.3f} ms", flush=True)
tl = torch.zeros(16, 64, dtype=torch.int64, device="cuda:0")
_lib.check_probe(lib.ll_probe_set_timeline(ptr(tl)))
for _ in range(2):
    run()
torch.cuda.synchronize()
_lib.check_probe(lib.ll_probe_set_timeline(None))
t = tl.cpu()
for i in range(2, 8):
    r = t[i]; z = int(r[0])
    g = lambda k: int(r[k]) - z
    print(f"tile {i}: [mma warp] tempty seen {g(3)}, windows staged seen {g(1)}, conv issued {g(2)}, norm mma {g(4)}..{g(5)} | "
          f"[epilogue warp 2] stage window {g(30)}..{g(31)}, tfull seen {g(33)}, E1 done {g(34)}, stage y^2 {g(36)}..{g(37)}, "
          f"E2 {g(50)}..{g(51)} | next tile start {int(t[i + 1][0]) - z}")
