"""SubbandAutoEncoderBerk timing: encode / decode of the level-0 subbands of 8 images, CTA-pair kernel vs single-CTA kernel."""
import functools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import SubbandAutoEncoderBerk
torch.manual_seed(0)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
orig = ops.igemm_tf32
with torch.no_grad():
    for ic, shape in ((3, (8, 3, 256, 384)), (1, (16, 1, 32, 48))):
        ae = SubbandAutoEncoderBerk(ic).to("cuda:0").eval()
        x = torch.randn(*shape, device="cuda:0")
        res = {}
        for pair in (True, False):
            ops.igemm_tf32 = functools.partial(orig, pair=pair)
            y = ae.encode(x)
            res[pair] = (y, ae.decode(y))
            print(f"in_ch={ic} {shape} pair={pair}: encode {t(lambda: ae.encode(x)):.2f} ms, decode {t(lambda: ae.decode(y)):.2f} ms")
        print("  pair == single:", torch.equal(res[True][0], res[False][0]), torch.equal(res[True][1], res[False][1]))
    ops.igemm_tf32 = orig
    # the two 3x3 GEMMs alone (8 x 256 x 384 px)
    for C, N in ((96, 192), (192, 96)):
        a = torch.randn(8, 256, 384, 2 * C, device="cuda:0")
        wp = ops.pack_tf32_weight(torch.randn(N, C, 3, 3, device="cuda:0") * 0.05)
        b = torch.zeros(N, device="cuda:0")
        for pair in (True, False):
            ms = t(lambda: ops.igemm_tf32(a, wp, b, N, epi=1, pair=pair), 10)
            fl = 2.0 * C * N * 9 * 8 * 256 * 384
            print(f"conv {C}->{N} pair={pair}: {ms:.3f} ms, useful {fl / ms / 1e9:.0f} TFLOP/s, issued {3 * fl / ms / 1e9:.0f} TFLOP/s")
    # fused conv + GDN kernels alone
    for C, N in ((96, 192), (192, 96)):
        a = torch.randn(8, 256, 384, 2 * C, device="cuda:0")
        wp = ops.pack_tf32_weight(torch.randn(N, C, 3, 3, device="cuda:0") * 0.05)
        gp = ops.pack_tf32_weight((torch.rand(N, N, device="cuda:0") * 0.01 + 0.1 * torch.eye(N, device="cuda:0")).reshape(N, N, 1, 1).contiguous())
        b = torch.zeros(N, device="cuda:0"); beta = torch.ones(N, device="cuda:0")
        ms = t(lambda: ops.igemm_tf32_gdn(a, wp, b, gp, beta, N), 10)
        fl = 2.0 * (C * 9 + N) * N * 8 * 256 * 384
        print(f"fused conv {C}->{N} + GDN: {ms:.3f} ms, useful {fl / ms / 1e9:.0f} TFLOP/s, issued {3 * fl / ms / 1e9:.0f} TFLOP/s")
