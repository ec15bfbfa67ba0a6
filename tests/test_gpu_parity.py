"""GPU parity tests (run with ``-m gpu`` on the B200 box): the CUDA path, called through the
C ABI (ctypes -> libll_b200.so), against (1) the committed golden vectors of the unmodified
reference and (2) the CPU oracle on seeded inputs, at sizes the oracle finishes in seconds.

Tolerances (north_star): quantised symbols bit-exact (every mismatch must be a one-step flip on
a rounding boundary of the oracle's pre-quantiser value -- ``flip_audit``); transform
coefficients and reconstructions within 1e-4 relative (||a-b||_inf / ||b||_inf) in fp32;
estimated bits per pixel within 0.1 %.
"""
import pytest
import torch

from oracle import lifting as olift, model as om, subband_ae as oae, thirdparty as tp

from common import bits_check, flip_audit, keyed_state, load_case, meta, product_model, rel_err

pytestmark = pytest.mark.gpu
META = meta()
EVAL_CASES = [k for k in META if k != "lifting_one_level" and not META[k]["training"]]
DEV = "cuda:0"


def _planes(model):
    return model.planes()


TC_LAYERS = ("conditioned2ZTsepSubbands",)   # entropy layers with a tensor-core (bf16) context path
PREC_CASES = [(k, "fp32") for k in EVAL_CASES] + \
             [(k, "bf16") for k in EVAL_CASES if META[k]["config"].get("entropy_layer") in TC_LAYERS]


@pytest.mark.parametrize("name,prec", PREC_CASES)
def test_model_forward_matches_reference_golden(name, prec):
    """fp32: every context CNN on the exact SIMT kernels (strict per-subband checks).  bf16 (the
    default): plc / cgp on the tcgen05 tensor cores -- symbols, coefficients and reconstruction are
    untouched (the context nets only feed the rate), bpp must stay within north_star's 0.1 %; the
    per-subband totals of these 16x16 .. 8x12 golden subbands get 0.5 % (BF16 operand rounding does
    not average out over a few hundred coefficients)."""
    m = META[name]
    model, cfg = product_model(dict(m["config"], ctx_precision=prec))
    tol_sub = 1e-3 if prec == "fp32" else 5e-3
    fo = 1e-3 if prec == "fp32" else 1.0
    keyed_state(model)
    model = model.to(DEV).eval()
    g = load_case(name)
    x = g["x"].to(DEV)
    acc = torch.zeros(1, dtype=torch.float64, device=DEV)
    model.set_bit_accumulator(acc)
    with torch.no_grad():
        xhat, si_xe, si_xo = model(x)
        # symbols and coefficients, plane by plane
        for c, sub in enumerate(_planes(model)):
            out_xe, out_xo = sub.autoencoder.encode(x if cfg.clrch == 3 else x[:, c:c + 1])
            _, _, xe_q, xo_q = sub.entropymodel(out_xe, out_xo)
            assert rel_err(out_xe.cpu(), g[f"out_xe_{c}"]) < 1e-4
            n, bad = flip_audit(xe_q.cpu(), g[f"xe_q_{c}"], g[f"out_xe_{c}"], label=f"{name}/{prec}/xe{c}")
            assert bad == 0 and n == 0, (n, bad)
            for i in range(cfg.dwtlevels):
                assert rel_err(out_xo[i].cpu(), g[f"out_xo_{c}_{i}"]) < 1e-4
                n, bad = flip_audit(xo_q[i].cpu(), g[f"xo_q_{c}_{i}"], g[f"out_xo_{c}_{i}"], label=f"{name}/{prec}/xo{c}_{i}")
                assert bad == 0 and n == 0, (c, i, n, bad)
    assert rel_err(xhat.cpu(), g["xhat"]) < 1e-4
    # (bf16: the coarsest-level causal chains -- si_xe and the last si_xo -- run their two dense layers on the tensor path too)
    assert bits_check(si_xe.cpu(), g["si_xe"], tol_sum=tol_sub, frac_outliers=fo)[2], bits_check(si_xe.cpu(), g["si_xe"])
    for i, s in enumerate(si_xo):
        assert bits_check(s.cpu(), g[f"si_xo_{i}"], tol_sum=tol_sub, frac_outliers=fo)[2], \
            (i, bits_check(s.cpu(), g[f"si_xo_{i}"]))
    B, _, H, W = g["x"].shape
    bits = float(si_xe.double().sum() + sum(s.double().sum() for s in si_xo))
    bpp = bits / (B * H * W)
    assert abs(bpp - m["bpp"]) <= 1e-3 * m["bpp"]          # bpp within 0.1 %
    # the in-kernel accumulator saw every subband twice (model() + the per-plane re-run above)
    assert abs(float(acc.item()) / 2 / (B * H * W) - m["bpp"]) <= 1e-3 * m["bpp"]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_training_mode_noise_parity(prec):
    """Training mode: noise is drawn in Python in the reference's call order and handed to the
    kernels; with the same noise tensors the rate matches the oracle."""
    name = "cdf97_cond2zt_L2_train"
    m = META[name]
    model, cfg = product_model(dict(m["config"], ctx_precision=prec))
    sd = keyed_state(model)
    model = model.to(DEV).train()
    g = load_case(name)
    # feed identical noise to both sides: CPU generator stream -> device tensors
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import compat
    torch.manual_seed(99)
    orig = compat.draw_noise
    compat.draw_noise = lambda like: torch.empty(like.shape, dtype=like.dtype).uniform_(-0.5, 0.5).to(like.device)
    try:
        with torch.no_grad():
            xhat, si_xe, si_xo = model(g["x"].to(DEV))
    finally:
        compat.draw_noise = orig
    assert rel_err(xhat.cpu(), g["xhat"]) < 1e-4
    assert bits_check(si_xe.cpu(), g["si_xe"], tol_sum=1e-3 if prec == "fp32" else 5e-3,
                      frac_outliers=1e-3 if prec == "fp32" else 1.0)[2], bits_check(si_xe.cpu(), g["si_xe"])
    tot, ref = float(si_xe.double().sum()), float(g["si_xe"].double().sum())
    for i, s in enumerate(si_xo):
        assert bits_check(s.cpu(), g[f"si_xo_{i}"], tol_sum=1e-3 if prec == "fp32" else 5e-3,
                          frac_outliers=1e-3 if prec == "fp32" else 1.0)[2]
        tot += float(s.double().sum())
        ref += float(g[f"si_xo_{i}"].double().sum())
    assert abs(tot - ref) <= 1e-3 * ref                    # bpp within 0.1 % in either precision


@pytest.mark.parametrize("mode", ["tc16", "tc", "fp32"])
def test_lifting_level_golden(mode):
    m = META["lifting_one_level"]
    model, cfg = product_model(dict(m["config"], lift_precision=mode))
    keyed_state(model)
    model = model.to(DEV).eval()
    g = load_case("lifting_one_level")
    ae = model.model0.autoencoder
    with torch.no_grad():
        LL, LH, HL, HH = ae.waveletForward[0].one_level_lifting(g["x"].to(DEV))
        rec = ae.waveletInverse[0].one_level_lifting(LL, LH, HL, HH)
    for a, k in ((LL, "LL"), (LH, "LH"), (HL, "HL"), (HH, "HH"), (rec, "rec")):
        assert rel_err(a.cpu(), g[k]) < 1e-5, k
    assert (rec.cpu() - g["x"]).abs().max().item() < 1e-5      # perfect reconstruction


@pytest.mark.parametrize("mode", ["tc16", "tc", "fp32"])
@pytest.mark.parametrize("shape,levels", [((2, 1, 48, 80), 3), ((1, 1, 128, 256), 4), ((3, 1, 16, 16), 2),
                                          ((1, 1, 64, 1024), 1), ((2, 1, 104, 106), 1)])
def test_learned_lifting_vs_oracle(shape, levels, mode):
    """Both arithmetic modes of the lifting kernels against the CPU oracle: "tc" = conv2/conv3 on tcgen05
    with the 3xTF32 split (default), "fp32" = every layer on the FP32 FMA pipe.  Same 1e-5 bar for both."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers import lifting_dwt_nets as ldn
    cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=levels,
                         lift_precision=mode)
    torch.manual_seed(1337)
    net = ldn.LiftingBasedNeuralWaveletv4(cfg)
    sd = om.keyed_weights({"m.autoencoder." + k: v for k, v in net.state_dict().items()})
    net.load_state_dict({k[len("m.autoencoder."):]: v for k, v in sd.items()}, strict=True)
    net = net.to(DEV).eval()
    torch.manual_seed(11)
    x = torch.rand(*shape) - 0.5
    with torch.no_grad():
        yl, yh = olift.transform_forward(x, sd, "m.autoencoder.", cfg)
        gl, gh = net.transform(x.to(DEV))
        assert rel_err(gl.cpu(), yl) < 1e-5
        for a, b in zip(gh, yh):
            assert rel_err(a.cpu(), b) < 1e-5
        rec = net.inverse_transform(gl, gh)
        assert rel_err(rec.cpu(), olift.transform_inverse(yl, yh, sd, "m.autoencoder.", cfg)) < 1e-5
        assert (rec.cpu() - x).abs().max().item() < 2e-5
        # stand-alone methods of the reference surface
        f0 = net.waveletForward[0]
        L, H = f0.lifting_forward_row_2_stage_lifting(x[:, :, 0::2].to(DEV), x[:, :, 1::2].to(DEV))
        oL, oH = olift.lift_rows_forward(x[:, :, 0::2], x[:, :, 1::2], sd, "m.autoencoder.waveletForward.0.", cfg)
        assert rel_err(L.cpu(), oL) < 1e-5 and rel_err(H.cpu(), oH) < 1e-5
        net.P_blocks[0].lift_precision = mode
        pb = net.P_blocks[0](x.to(DEV))
        assert rel_err(pb.cpu(), olift.p_block(x, sd, "m.autoencoder.P_blocks.0.")) < 1e-5


@pytest.mark.parametrize("shape,J", [((2, 3, 64, 96), 3), ((1, 1, 16, 24), 3), ((1, 2, 8, 8), 2), ((1, 1, 256, 256), 4),
                                     ((2, 1, 4, 36), 1)])
def test_dwt97_vs_oracle(shape, J):
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    torch.manual_seed(5)
    x = torch.rand(*shape) - 0.5
    yl, yh = tp.dwt97_forward(x, J)
    gl, gh = ops.dwt97_forward(x.to(DEV), J)
    assert rel_err(gl.cpu(), yl) < 1e-5
    for a, b in zip(gh, yh):
        assert rel_err(a.cpu(), b) < 1e-5
    rec = ops.dwt97_inverse(gl, gh)
    assert rel_err(rec.cpu(), tp.dwt97_inverse(yl, yh)) < 1e-5
    if min(shape[2], shape[3]) >> (J - 1) >= 10:   # true periodisation => perfect reconstruction
        assert (rec.cpu() - x).abs().max().item() < 1e-5


def test_dwt97_inverse_output_alignment_paths():
    """The fast inverse kernel writes an item's 8 outputs with one 256-bit store when the output plane is
    32-byte aligned and with two 128-bit stores otherwise: both paths through the C ABI (ll_dwt97_inv_level,
    include/ll_api.h) on the same subbands must give the same bits, and match the oracle."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib, ops
    lib = _lib.load()
    torch.manual_seed(11)
    N, h, w = 3, 64, 160
    x = torch.rand(N, 1, h, w) - 0.5
    yl, yh = tp.dwt97_forward(x, 1)
    want = tp.dwt97_inverse(yl, yh).reshape(N, h, w)
    gl, gh = yl.to(DEV).contiguous(), yh[0].to(DEV).contiguous()
    sub = (h // 2) * (w // 2)
    outs = []
    for off in (0, 4):                       # floats: 0 -> 32-byte aligned base, 4 -> 16-byte aligned only
        buf = torch.full((N * h * w + 8,), float("nan"), device=DEV)
        out = buf[off:off + N * h * w]
        assert out.data_ptr() % 32 == (16 if off else 0)
        with torch.cuda.device(DEV):
            ops.check(lib.ll_dwt97_inv_level(ops.ptr(gl), sub, ops.ptr(gh), 3 * sub, out.data_ptr(), h * w, N, h, w,
                                             ops.stream_ptr()))
        torch.cuda.synchronize()
        assert torch.isnan(buf[:off]).all() and torch.isnan(buf[off + N * h * w:]).all()   # nothing outside the plane
        outs.append(out.view(N, h, w).cpu())
    assert torch.equal(outs[0], outs[1])
    assert rel_err(outs[0], want) < 1e-5


def test_pointwise_autoencoder_symbols_bit_exact():
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import SubbandAutoEncoder
    torch.manual_seed(1337)
    ae = SubbandAutoEncoder(3)
    sd = om.keyed_weights({"autoencoder.Yh_ae.0." + k: v for k, v in ae.state_dict().items()})
    ae.load_state_dict({k[len("autoencoder.Yh_ae.0."):]: v for k, v in sd.items()})
    ae = ae.to(DEV)
    x = torch.randn(2, 3, 37, 53) * 0.7
    with torch.no_grad():
        oy = oae.encode(x, sd, "autoencoder.Yh_ae.0.")
        gy, gq = ae.encode_and_round(x.to(DEV))
        assert rel_err(gy.cpu(), oy) < 1e-5
        n, bad = flip_audit(gq.cpu(), torch.round(oy), oy)
        assert bad == 0 and n <= 2
        od = oae.decode(torch.round(oy), sd, "autoencoder.Yh_ae.0.")
        assert rel_err(ae.decode(torch.round(oy).to(DEV)).cpu(), od) < 1e-5


def test_edge_cases_and_errors():
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import _lib, ops
    # empty batch: nothing to do, no error
    gl, gh = ops.dwt97_forward(torch.empty(0, 3, 16, 16, device=DEV), 2)
    assert gl.shape == (0, 3, 4, 4) and gh[0].shape == (0, 3, 3, 8, 8)
    # H, W not divisible by 2^L -> Python error like the reference's shape errors
    with pytest.raises(ValueError):
        ops.dwt97_forward(torch.zeros(1, 1, 20, 16, device=DEV), 3)
    # the C ABI itself rejects odd extents with LL_EINVAL and a message
    lib = _lib.load()
    x = torch.zeros(1, 1, 15, 16, device=DEV)
    rc = lib.ll_dwt97_fwd_level(x.data_ptr(), 240, x.data_ptr(), 60, x.data_ptr(), 180, 1, 15, 16, None)
    assert rc == _lib.LL_EINVAL and b"even" in lib.ll_last_error()
    # CPU tensors are refused: there is no fallback
    with pytest.raises(RuntimeError):
        ops.dwt97_forward(torch.zeros(1, 1, 16, 16), 1)


def test_full_size_properties():
    """BASELINE config-2 size (one colour plane, batch 2 to bound the time): size-independent
    properties -- perfect reconstruction of the 4-level learned lifting and batch independence."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers import lifting_dwt_nets as ldn
    cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", dwtlevels=4)
    torch.manual_seed(1337)
    net = ldn.LiftingBasedNeuralWaveletv4(cfg).to(DEV).eval()
    torch.manual_seed(1)
    x = (torch.rand(2, 1, 512, 768, device=DEV) - 0.5)
    with torch.no_grad():
        yl, yh = net.transform(x)
        rec = net.inverse_transform(yl, yh)
        assert (rec - x).abs().max().item() < 2e-5
        yl1, yh1 = net.transform(x[1:2])
        assert torch.equal(yl1, yl[1:2]) and all(torch.equal(a, b[1:2]) for a, b in zip(yh1, yh))


def test_full_size_codec_properties():
    """BASELINE config-3 image size (512x768, batch 1): size-independent properties of the whole codec forward --
    (1) the tensor-core context path moves bpp by < 0.1 % relative to the exact-fp32 context path and changes no
    symbol and no reconstructed sample (the context CNNs only feed the rate); (2) both lifting arithmetic modes give
    the same symbols up to rounding-boundary flips and the same reconstruction to 1e-4."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
        LiftingBasedDWTNetWrapper
    torch.manual_seed(2)
    x = om.preprocess(torch.rand(1, 3, 512, 768)).to(DEV)
    outs = {}
    for tag, kw in (("tc_bf16", {}), ("tc_fp32ctx", dict(ctx_precision="fp32")), ("fp32_lift", dict(lift_precision="fp32"))):
        cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder",
                             entropy_layer="conditioned2ZTsepSubbands", dwtlevels=4, **kw)
        torch.manual_seed(1337)
        model = LiftingBasedDWTNetWrapper(cfg)
        keyed_state(model)
        model = model.to(DEV).eval()
        with torch.no_grad():
            xhat, si_xe, si_xo = model(x)
            sub = model.model0
            oxe, oxo = sub.autoencoder.encode(x[:, 0:1])
            _, _, xe_q, xo_q = sub.entropymodel(oxe, oxo)
            dec0 = sub.autoencoder.decode(xe_q, list(xo_q))
            if tag == "fp32_lift":          # the fp32-lifting symbols decoded by the default (tensor-core) inverse transform
                forced = outs["tc_bf16"][4].autoencoder.decode(xe_q, list(xo_q))
        bits = float(si_xe.double().sum() + sum(s.double().sum() for s in si_xo))
        outs[tag] = (xhat, bits, [xe_q] + list(xo_q), [oxe] + list(oxo), sub, dec0)
        del model
    a, b, c = outs["tc_bf16"], outs["tc_fp32ctx"], outs["fp32_lift"]
    assert abs(a[1] - b[1]) <= 1e-3 * b[1]                                   # bpp within 0.1 %
    assert torch.equal(a[0], b[0]) and all(torch.equal(p, q) for p, q in zip(a[2], b[2]))
    flips = 0
    for qa, qc, pre in zip(a[2], c[2], c[3]):
        n, bad = flip_audit(qa.cpu(), qc.cpu(), pre.cpu())
        assert bad == 0 and n <= max(2, qa.numel() // 20000), (n, bad)     # rounding-boundary flips only
        flips += n
    # reconstruction: on IDENTICAL symbols the two lifting arithmetics agree to 1e-4 (a flipped symbol moves the samples
    # under its synthesis footprint by a quantisation step, so whole reconstructions are only compared when nothing flipped)
    assert rel_err(forced, c[5]) < 1e-4, rel_err(forced, c[5])
    if flips == 0:
        assert rel_err(a[0], c[0]) < 1e-4


def test_config5_tile_five_levels_vs_oracle():
    """BASELINE config 5: a 2048x2048 image cut into 8 independent (3, 2048, 256) tiles, 5-level learned lifting +
    conditioned2ZT.  One tile, one colour plane's networks: (1) the 5-level transform against the CPU oracle (coefficients
    1e-5, reconstruction 2e-5); (2) the whole codec forward on the tile: symbols bit-exact up to rounding-boundary flips,
    bits within 0.1 % (exact-fp32 context path), reconstruction 1e-4."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import parallel
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
        LiftingBasedDWTNetWrapper
    boxes = parallel.tiles_of(2048, 2048, 8)
    y0, y1, x0, x1 = boxes[3]
    assert ((y1 - y0) % 32 == 0) and ((x1 - x0) % 32 == 0)
    torch.manual_seed(5)
    img = om.preprocess(torch.rand(1, 3, 2048, 2048))
    tile = img[:, :, y0:y1, x0:x1].contiguous()
    cfg = om.default_cfg(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder",
                         entropy_layer="conditioned2ZTsepSubbands", dwtlevels=5, ctx_precision="fp32")
    torch.manual_seed(1337)
    model = LiftingBasedDWTNetWrapper(cfg)
    sd = keyed_state(model)
    model = model.to(DEV).eval()
    x = tile[:, 0:1]
    with torch.no_grad():
        pre = "model0.autoencoder."
        yl, yh = olift.transform_forward(x, sd, pre, cfg)
        ae = model.model0.autoencoder
        gl, gh = ae.transform(x.to(DEV))
        assert len(gh) == 5 and gl.shape[-2:] == ((y1 - y0) // 32, (x1 - x0) // 32)
        assert rel_err(gl.cpu(), yl) < 1e-5
        for a, b in zip(gh, yh):
            assert rel_err(a.cpu(), b) < 1e-5
        rec = ae.inverse_transform(gl, gh)
        assert (rec.cpu() - x).abs().max().item() < 2e-5
        # whole codec forward of this plane on the tile against the oracle
        oxhat, osi_xe, osi_xo, oxe_q, oxo_q, o_xe, o_xo = om.net_forward(x, sd, "model0.", cfg)
        oxe, oxo = ae.encode(x.to(DEV))
        si_xe, si_xo, xe_q, xo_q = model.model0.entropymodel(oxe, oxo)
        flips = 0
        for q, oq, pre_q in zip([xe_q] + list(xo_q), [oxe_q] + list(oxo_q), [o_xe] + list(o_xo)):
            n, bad = flip_audit(q.cpu(), oq, pre_q)
            assert bad == 0 and n <= max(2, q.numel() // 20000), (n, bad)
            flips += n
        # decoder parity on the oracle's own symbols (a rounding-boundary flip legitimately moves the reconstruction
        # by a quantisation step), and the end-to-end reconstruction when no symbol flipped
        xhat = ae.decode(oxe_q.to(DEV), [t.to(DEV) for t in oxo_q])
        if flips == 0:
            assert rel_err(ae.decode(xe_q, xo_q).cpu(), oxhat) < 1e-4
        bits = float(si_xe.double().sum() + sum(t.double().sum() for t in si_xo))
        obits = float(osi_xe.double().sum() + sum(t.double().sum() for t in osi_xo))
        assert abs(bits - obits) <= 1e-3 * obits
        assert rel_err(xhat.cpu(), oxhat) < 1e-4


@pytest.mark.parametrize("entropy_layer", ["conditioned2ZTsepSubbands", "onlyEZWT"])
def test_cuda_graph_replay_matches_eager(entropy_layer):
    """utils.cuda_graph.GraphedForward: the whole codec forward (three planes on three streams, every launch through
    the C ABI) captured once and replayed must reproduce the eager call bit for bit, also on fresh input data."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils.cuda_graph import GraphedForward
    cfg = dict(netType="LiftingBasedNeuralWaveletv4", autoencoder="SubbandAutoEncoder", entropy_layer=entropy_layer,
               dwtlevels=3, clrch=1)
    model, _ = product_model(cfg)
    model = model.to(DEV).eval()
    torch.manual_seed(3)
    x0 = om.preprocess(torch.rand(2, 3, 64, 96)).to(DEV)
    x1 = om.preprocess(torch.rand(2, 3, 64, 96)).to(DEV)
    graphed = GraphedForward(model, x0)
    for x in (x0, x1, x0):
        with torch.no_grad():
            want = model(x)
        got = graphed(x)
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        assert len(got[2]) == len(want[2]) and all(torch.equal(a, b) for a, b in zip(got[2], want[2]))
    with pytest.raises(ValueError):
        graphed(x0[:1])


@pytest.mark.parametrize("shape", [(2, 32, 17, 23), (1, 32, 1, 1), (3, 32, 64, 96), (0, 32, 8, 8)])
def test_pointwise_mlp_tail_matches_torch(shape):
    """ll_pw_mlp3 (the fused 1x1 tail of the ZTBlock dependency CNNs, LiftingBasedDWT_net.py:618-680) against the same
    three convs evaluated by torch in float64: fp32 rounding only (1e-5 relative), odd plane sizes included."""
    import torch.nn as nn
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    torch.manual_seed(21)
    convs = [nn.Conv2d(32, 32, 1), nn.Conv2d(32, 32, 1), nn.Conv2d(32, 1, 1)]
    x = torch.randn(*shape)
    with torch.no_grad():
        ref = x.double()
        for i, c in enumerate(convs):
            ref = nn.functional.conv2d(ref, c.weight.double(), c.bias.double())
            if i < 2:
                ref = nn.functional.leaky_relu(ref, 0.01)
    got = ops.pw_mlp3(x.to(DEV), *[c.to(DEV) for c in convs])
    assert got.shape == (shape[0], 1, shape[2], shape[3])
    if shape[0]:
        assert rel_err(got.cpu().double(), ref) < 1e-5
    with pytest.raises(ValueError):
        ops.pw_mlp3(torch.zeros(1, 16, 4, 4, device=DEV), *[c.to(DEV) for c in convs])
