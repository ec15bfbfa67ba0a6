"""Per-tile timeline of igemm_tf32_gdn_pair_kernel (probe build with LL_TIMELINE, csrc/probe/igemm_timeline.cu): SM cycles
between the hand-offs of CTA 0's MMA thread and of one epilogue warp, for the two large instances of the scaling network."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib, ops
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.ops import ptr, stream_ptr
lib = _lib.load_probe()
P = ctypes.c_void_p
lib.ll_igemm_tf32_gdn.restype = ctypes.c_int
lib.ll_igemm_tf32_gdn.argtypes = [P] * 5 + [ctypes.c_int] * 7 + [P, P]
def timed(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
torch.manual_seed(0)
for C, N in ((96, 192), (192, 96)):
    a = torch.randn(8, 256, 384, 2 * C, device="cuda:0")
    wp = ops.pack_tf32_weight(torch.randn(N, C, 3, 3, device="cuda:0") * 0.05)
    gp = ops.pack_tf32_weight((torch.rand(N, N, device="cuda:0") * 0.01 + 0.1 * torch.eye(N, device="cuda:0")).reshape(N, N, 1, 1).contiguous())
    b = torch.zeros(N, device="cuda:0"); beta = torch.ones(N, device="cuda:0")
    sz = torch.empty(8, 256, 384, 2 * N, device="cuda:0")
    run = lambda: _lib.check_probe(lib.ll_igemm_tf32_gdn(ptr(a), ptr(wp), ptr(b), ptr(gp), ptr(beta), 8, 256, 384, C, N, 9, 0, ptr(sz), stream_ptr()))
    for ns in (1, 0):
        lib.ll_probe_set_nostore(ns)
        print(f"conv {C}->{N} + GDN, E2 global stores {'off' if ns else 'on'}: {timed(run):.3f} ms", flush=True)
    tl = torch.zeros(16, 64, dtype=torch.int64, device="cuda:0")
    _lib.check_probe(lib.ll_probe_set_timeline(ptr(tl)))
    for _ in range(2):
        _lib.check_probe(lib.ll_igemm_tf32_gdn(ptr(a), ptr(wp), ptr(b), ptr(gp), ptr(beta), 8, 256, 384, C, N, 9, 0, ptr(sz), stream_ptr()))
    torch.cuda.synchronize()
    _lib.check_probe(lib.ll_probe_set_timeline(None))
    t = tl.cpu()
    print(f"== conv {C}->{N} + GDN: cycles relative to the tile's mainloop start (slot 1)")
    if int(t[9][60]) and int(t[2][60]):
        print(f"SM clock under load: {(int(t[9][1]) - int(t[2][1])) / (int(t[9][60]) - int(t[2][60])):.3f} GHz (tiles 2..9 of CTA 0)")
    rounds = 6 if N == 192 else 1
    passes = 2 if N == 192 else 1
    for i in range(2, 8):
        r = t[i]; z = int(r[1])
        g = lambda k: int(r[k]) - z
        nxt = int(t[i + 1][1]) - z
        s = f"tile {i}: wait_tempty {int(r[1]) - int(r[0])}, mainloop issued {g(2)}, tfull seen {g(33)}, E1 done {g(34)}"
        for k in range(rounds):
            s += f" | r{k}: stage {g(36 + 2 * k)}..{g(37 + 2 * k)} mma {g(4 + 2 * k)}..{g(5 + 2 * k)}"
        for ps in range(passes):
            s += f" | E2[{ps}] {g(50 + 2 * ps)}..{g(51 + 2 * ps)}"
        s += f" | next mainloop start {nxt}"
        print(s)
