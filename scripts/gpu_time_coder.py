import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import model as om
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import onlyEZWT
dev = "cuda:0"; B, H, W, L = 16, 512, 768, 4
cfg = om.default_cfg(entropy_layer="onlyEZWT", dwtlevels=L)
torch.manual_seed(1337); em = onlyEZWT(cfg).to(dev).eval()
xe = torch.randn(B, 1, H >> L, W >> L, device=dev) * 4
xo = [torch.randn(B, 3, H >> (l + 1), W >> (l + 1), device=dev) * (1.5 + l) for l in range(L)]
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): o = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, o
with torch.no_grad():
    print("forward %.1f ms" % t(lambda: em(xe, xo))[0])
    ms_c, (streams, xe_q, qs) = t(lambda: em.compress(xe, xo)); print("compress %.1f ms" % ms_c)
    print("decompress %.1f ms" % t(lambda: em.decompress(streams))[0])
    mss = []; em(xe, xo, keep_ms=mss)
    for target in (8192, 2048, 512):
        S = ops.rans_streams_per_image(qs[0][0].numel(), target)
        e, (w, c, S) = t(lambda: ops.rans_encode(ops.RANS_GAUSS, qs[0], mss[0], S))
        d, back = t(lambda: ops.rans_decode(ops.RANS_GAUSS, w, c, mss[0], qs[0].shape, S))
        print(f"level-0 tensor {tuple(qs[0].shape)}: {S} streams/image: encode {e:.2f} ms, decode {d:.2f} ms, {16.0 * w.numel() / qs[0].numel():.3f} bits/sample (+{64.0 * c.numel() / qs[0].numel():.4f} stream overhead), exact {torch.equal(back, qs[0])}")
