"""GPU tests of the tcgen05 implicit-GEMM context-CNN kernels (``-m gpu``).

These are floating-point kernels with BF16 operands and FP32 accumulation, so the reference is a
plain torch fp32 convolution of the *same bf16-rounded operands* (then the difference is only the
summation order: tolerance 2e-3 of the output scale), plus a looser check against the un-rounded
fp32 convolution (BF16 operand rounding: 2^-9 relative per operand).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
    return ops


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _nhwc_bf16(x, cpad):
    B, C, H, W = x.shape
    out = torch.zeros(B, H, W, cpad, dtype=torch.bfloat16, device=x.device)
    out[..., :C] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out.contiguous()


@pytest.mark.parametrize("B,Cin,Cout,H,W,taps,lrelu", [
    (1, 64, 16, 8, 16, 1, False),        # one tile, one k-block, smallest N
    (1, 64, 64, 8, 16, 9, False),        # 3x3 halo through TMA zero fill
    (2, 243, 243, 24, 40, 9, False),     # plc shape: padded K and N, partial tiles in x
    (2, 162, 162, 16, 24, 1, True),      # cgp layer: 1x1, Npad = 176, LeakyReLU
    (1, 243, 243, 4, 6, 9, False),       # plane smaller than the TMA box
    (3, 128, 256, 40, 48, 9, True),      # more tiles than accumulator stages per CTA? (45 tiles, 148 CTAs: no)
    (8, 256, 256, 64, 96, 9, False),     # 384 tiles: several tiles per CTA, pipeline wrap-around
])
def test_igemm_conv_matches_torch(B, Cin, Cout, H, W, taps, lrelu):
    ops = _ops()
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + Cin + H)
    k = 3 if taps == 9 else 1
    x = (torch.rand(B, Cin, H, W, generator=g) * 2 - 1).to(DEV)
    w = ((torch.rand(Cout, Cin, k, k, generator=g) * 2 - 1) * (3.0 / (Cin * taps)) ** 0.5).to(DEV)
    b = (torch.rand(Cout, generator=g) - 0.5).to(DEV)
    kpad = (Cin + 63) // 64 * 64
    xn = _nhwc_bf16(x, kpad)
    wp = ops.pack_igemm_weight(w)
    y = ops.igemm_conv(xn, wp, b, Cout, lrelu=lrelu)
    ref = F.conv2d(_bf(x).double(), _bf(w).double(), b.double(), padding=k // 2)
    ref32 = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2)
    if lrelu:
        ref, ref32 = F.leaky_relu(ref, 0.01), F.leaky_relu(ref32, 0.01)
    scale = ref.abs().max().item()
    assert (y.double() - ref).abs().max().item() <= 2e-3 * scale        # same operands, summation order only
    assert (y.double() - ref32).abs().max().item() <= 2e-2 * scale      # bf16 operand rounding


def test_igemm_outputs_remap_and_nhwc():
    """fp32 NCHW output through the (plc0,csc0,plc1,csc1,plc2,csc2) channel remap and the bf16 NHWC
    output at a channel offset, both from one launch."""
    ops = _ops()
    torch.manual_seed(3)
    B, Cin, Cout, H, W = 2, 243, 243, 16, 32
    x = (torch.rand(B, Cin, H, W) * 2 - 1).to(DEV)
    w = ((torch.rand(Cout, Cin, 3, 3) * 2 - 1) * 0.04).to(DEV)
    b = (torch.rand(Cout) - 0.5).to(DEV)
    xn = _nhwc_bf16(x, 256)
    wp = ops.pack_igemm_weight(w)
    cat = torch.full((B, 486, H, W), 7.0, device=DEV)
    nh = torch.full((B, H, W, 576), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.igemm_conv(xn, wp, b, Cout, out=cat, co_group=81, co_stride=162, co_off=0, out_nhwc=nh, nhwc_coff=192)
    npad = wp.shape[1]                                                   # 256: the NHWC path writes every padded channel
    ref = F.conv2d(_bf(x), _bf(w), b, padding=1)
    scale = ref.abs().max().item()
    for k in range(3):
        got = cat[:, 162 * k:162 * k + 81]
        assert (got - ref[:, 81 * k:81 * k + 81]).abs().max().item() <= 2e-3 * scale
        assert (cat[:, 162 * k + 81:162 * k + 162] == 7.0).all()       # csc slots untouched
    got = nh[..., 192:192 + 243].float().permute(0, 3, 1, 2)
    assert (got - ref).abs().max().item() <= 1e-2 * scale               # bf16 output rounding
    assert (nh[..., :192] == 7.0).all() and (nh[..., 192 + 243:192 + npad] == 0).all() and (nh[..., 192 + npad:] == 7.0).all()


@pytest.mark.parametrize("up", [True, False])
def test_ctx_conv_nhwc_head_matches_torch(up):
    ops = _ops()
    torch.manual_seed(4)
    B, H, W = 2, 12, 20
    con = torch.round(torch.randn(B, 3, H // 2 if up else H, W // 2 if up else W) * 3).to(DEV)
    w = ((torch.rand(243, 3, 3, 3) * 2 - 1) * 0.3).to(DEV)
    b = (torch.rand(243) - 0.5).to(DEV)
    out = ops.ctx_conv_nhwc(con, w, b, upsample2=up, lrelu=True, region=256)
    assert tuple(out.shape) == (B, H, W, 256) and out.dtype == torch.bfloat16
    src = con.repeat_interleave(2, 2).repeat_interleave(2, 3) if up else con
    ref = F.leaky_relu(F.conv2d(src, w, b, padding=1), 0.01)
    got = out[..., :243].float().permute(0, 3, 1, 2)
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert (out[..., 243:] == 0).all()


def test_ctx_conv_nhwc_masked_grouped_csc():
    """csc_list[i]: MaskedConv2d('A', 3, 243, 5, padding=2, groups=3) written into 128-channel group slots."""
    ops = _ops()
    torch.manual_seed(6)
    B, H, W = 2, 10, 18
    q = torch.round(torch.randn(B, 3, H, W) * 4).to(DEV)
    w = ((torch.rand(243, 1, 5, 5) * 2 - 1) * 0.2)
    mask = torch.ones(5, 5)
    mask[2, 2:] = 0
    mask[3:] = 0
    w = (w * mask).to(DEV)
    b = (torch.rand(243) - 0.5).to(DEV)
    out = torch.full((B, H, W, 256 + 384), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.ctx_conv_nhwc(q, w, b, groups=3, live_taps=12, out=out, coff=256, co_group=81, co_gstride=128, region=384)
    ref = F.conv2d(q, w, b, padding=2, groups=3)
    for g in range(3):
        got = out[..., 256 + 128 * g:256 + 128 * g + 81].float().permute(0, 3, 1, 2)
        assert (got - ref[:, 81 * g:81 * g + 81]).abs().max().item() <= 1e-2 * ref.abs().max().item()
        assert (out[..., 256 + 128 * g + 81:256 + 128 * (g + 1)] == 0).all()
    assert (out[..., :256] == 7.0).all()


def test_ctx_im2col_split_gemm_matches_fp32_convs():
    """The plc head and the masked csc as 1-tap tensor-core GEMMs over ll_ctx_im2col's split rows: values far beyond bf16's
    8 bits (quantised coefficients) must come out as accurately as from the fp32 SIMT convs (both round the OUTPUT to bf16)."""
    ops = _ops()
    torch.manual_seed(14)
    B, H, W = 2, 12, 20
    con = (torch.round(torch.randn(B, 3, H // 2, W // 2) * 300) + torch.rand(B, 3, H // 2, W // 2) * 0.37).to(DEV)   # round(x - mu) + mu
    q = (torch.round(torch.randn(B, 3, H, W) * 200) + 0.123).to(DEV)
    wh = ((torch.rand(243, 3, 3, 3) * 2 - 1) * 0.3).to(DEV)
    bh = (torch.rand(243) - 0.5).to(DEV)
    wc = ((torch.rand(243, 1, 5, 5) * 2 - 1) * 0.2)
    mask = torch.ones(5, 5)
    mask[2, 2:] = 0
    mask[3:] = 0
    wc = (wc * mask).to(DEV)
    bc = (torch.rand(243) - 0.5).to(DEV)
    a = ops.ctx_im2col(con, q)
    assert tuple(a.shape) == (B, H, W, 320) and a.dtype == torch.bfloat16
    # the rows themselves: hi + lo reproduces the fp32 window to 16 bits, unused channels are zero
    up = con.repeat_interleave(2, 2).repeat_interleave(2, 3)
    cols = F.unfold(up, 3, padding=1).reshape(B, 27, H, W).permute(0, 2, 3, 1)
    af = a.float()
    assert torch.equal(af[..., 0:27], af[..., 54:81])
    assert ((af[..., 0:27] + af[..., 27:54]) - cols).abs().max().item() <= 2.0 ** -15 * cols.abs().max().item()
    assert (a[..., 81:128] == 0).all() and (a[..., 128 + 36:192] == 0).all()
    whp = torch.zeros(243, 128, 1, 1, device=DEV)
    whp[:, :81, 0, 0] = ops.split_bf16_weight(wh.reshape(243, 27))
    t = torch.full((B, H, W, 256), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.igemm_conv(a, ops.pack_igemm_weight(whp, npad=256, kpad=128), bh, 243, lrelu=True, out_nhwc=t, koff=[[0, 64]])
    ref = F.leaky_relu(F.conv2d(up, wh, bh, padding=1), 0.01).permute(0, 2, 3, 1)
    old = ops.ctx_conv_nhwc(con, wh, bh, upsample2=True, lrelu=True, region=256)
    scale = ref.abs().max().item()
    assert (t[..., :243].float() - ref).abs().max().item() <= 2.0 ** -8 * scale
    assert (t[..., :243].float() - old[..., :243].float()).abs().max().item() <= 2.0 ** -7 * scale   # one bf16 ulp apart at most
    assert (t[..., 243:] == 0).all()
    wcp = []
    for g in range(3):
        wg = torch.zeros(81, 64, 1, 1, device=DEV)
        wg[:, :36, 0, 0] = ops.split_bf16_weight(wc[81 * g:81 * (g + 1), 0].reshape(81, 25)[:, :12])
        wcp.append(ops.pack_igemm_weight(wg, npad=128, kpad=64))
    g_in = torch.full((B, H, W, 640), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.igemm_conv(a, torch.stack(wcp).contiguous(), bc, 81, out_nhwc=g_in, nhwc_coff=256, nhwc_gstride=128,
                   koff=[[128], [192], [256]])
    refc = F.conv2d(q, wc, bc, padding=2, groups=3).permute(0, 2, 3, 1)
    sc = refc.abs().max().item()
    for g in range(3):
        got = g_in[..., 256 + 128 * g:256 + 128 * g + 81].float()
        assert (got - refc[..., 81 * g:81 * g + 81]).abs().max().item() <= 2.0 ** -8 * sc
        assert (g_in[..., 256 + 128 * g + 81:256 + 128 * (g + 1)] == 0).all()
    assert (g_in[..., :256] == 7.0).all()


def test_igemm_grouped_1x1_with_koff():
    """cgp layer shape: 3 groups, each reading 4 k-blocks at per-group channel offsets of one NHWC tensor."""
    ops = _ops()
    torch.manual_seed(8)
    B, H, W, C = 2, 16, 24, 640
    x = (torch.rand(B, H, W, C) * 2 - 1).to(torch.bfloat16).to(DEV)
    koff = [[0, 64, 256, 320], [64, 128, 384, 448], [128, 192, 512, 576]]
    wg = [((torch.rand(162, 256, 1, 1) * 2 - 1) * 0.1).to(DEV) for _ in range(3)]
    bias = (torch.rand(3 * 162) - 0.5).to(DEV)
    wp = torch.stack([ops.pack_igemm_weight(w, npad=192, kpad=256) for w in wg]).contiguous()
    h1 = torch.full((B, H, W, 576), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.igemm_conv(x, wp, bias, 162, lrelu=True, out_nhwc=h1, nhwc_gstride=192, koff=koff)
    f32 = ops.igemm_conv(x, wp, bias, 162, lrelu=True, koff=koff)          # fp32 NCHW, channel g*162 + c
    xf = x.float()
    for g in range(3):
        xin = torch.cat([xf[..., o:o + 64] for o in koff[g]], dim=-1)       # (B,H,W,256)
        ref = F.leaky_relu(xin @ _bf(wg[g][:, :, 0, 0]).t() + bias[162 * g:162 * (g + 1)], 0.01)
        scale = ref.abs().max().item()
        assert (f32[:, 162 * g:162 * (g + 1)].permute(0, 2, 3, 1) - ref).abs().max().item() <= 2e-3 * scale
        assert (h1[..., 192 * g:192 * g + 162].float() - ref).abs().max().item() <= 1e-2 * scale
        assert (h1[..., 192 * g + 162:192 * (g + 1)] == 0).all()             # padded channels are exact zeros


def test_cgp_tail_rate_matches_torch():
    ops = _ops()
    from oracle import thirdparty as tp
    torch.manual_seed(9)
    B, G, H, W = 2, 3, 12, 20
    h2 = torch.randn(B, G * 54, H, W)
    w3, b3 = torch.randn(G * 18, 54, 1, 1) * 0.2, torch.randn(G * 18) * 0.1
    w4, b4 = torch.randn(G * 2, 18, 1, 1) * 0.3, torch.randn(G * 2) * 0.1
    b4[0::2] += 3.0
    x = torch.randn(B, G, H, W) * 5
    a = F.leaky_relu(F.conv2d(h2, w3, b3, groups=G), 0.01)
    ms = F.conv2d(a, w4, b4, groups=G)
    _, lik = tp.gaussian_conditional_forward(x, ms[:, 0::2], ms[:, 1::2], False)
    ref = -torch.log2(lik)
    acc = torch.zeros(1, dtype=torch.float64, device=DEV)
    bits, gms = ops.cgp_tail_rate(h2.to(DEV), w3.to(DEV), b3.to(DEV), w4.to(DEV), b4.to(DEV), x.to(DEV), want_ms=True, acc=acc)
    assert (gms.cpu() - ms).abs().max().item() <= 1e-4 * ms.abs().max().item()
    rs = abs(float(bits.double().sum().cpu() - ref.double().sum())) / float(ref.double().sum())
    assert rs < 1e-3
    assert abs(float(acc.item()) - float(bits.double().sum().item())) <= 1e-3 * float(acc.item())


@pytest.mark.parametrize("B,H,W,train", [(2, 12, 20, False), (1, 37, 50, False), (3, 16, 24, True)])
def test_fused_cgp_layer2_tail_rate_equals_the_two_kernel_chain(B, H, W, train):
    """ll_igemm_cgp_tail (cgp layer 2 as a grouped 1x1 tcgen05 GEMM, layers 3-4 and the Gaussian rate in its epilogue) against
    ll_igemm_conv + ll_cgp_tail_rate on the same operands: the same products in the same order -> identical bits."""
    ops = _ops()
    torch.manual_seed(21)
    G = 3
    h1 = (torch.rand(B, H, W, 576) * 2 - 1).to(torch.bfloat16).to(DEV)
    k2 = [[192 * g, 192 * g + 64, 192 * g + 128] for g in range(G)]
    w2 = [((torch.rand(54, 162, 1, 1) * 2 - 1) * 0.15) for _ in range(G)]
    for g in range(G):
        w2[g] = torch.nn.functional.pad(w2[g], (0, 0, 0, 0, 0, 30)).to(DEV)      # K 162 -> 192
    wp2 = torch.stack([ops.pack_igemm_weight(w, npad=64, kpad=192) for w in w2]).contiguous()
    b2 = (torch.rand(G * 54) - 0.5).to(DEV)
    w3, b3 = (torch.randn(G * 18, 54, 1, 1) * 0.2).to(DEV), (torch.randn(G * 18) * 0.1).to(DEV)
    w4, b4 = (torch.randn(G * 2, 18, 1, 1) * 0.3).to(DEV), torch.randn(G * 2) * 0.1
    b4[0::2] += 3.0
    b4 = b4.to(DEV)
    x = (torch.randn(B, G, H, W) * 5).to(DEV)
    noise = (torch.rand(B, G, H, W) - 0.5).to(DEV) if train else None
    acc_a = torch.zeros(1, dtype=torch.float64, device=DEV)
    acc_b = torch.zeros(1, dtype=torch.float64, device=DEV)
    h2 = ops.igemm_conv(h1, wp2, b2, 54, lrelu=True, koff=k2)
    ref = ops.cgp_tail_rate(h2, w3, b3, w4, b4, x, noise, acc=acc_a)
    got = ops.igemm_cgp_tail(h1, wp2, b2, 54, k2, w3, b3, w4, b4, x, noise, acc=acc_b)
    assert torch.isfinite(got).all()
    assert torch.equal(got, ref)
    assert abs(acc_a.item() - acc_b.item()) <= 1e-5 * abs(acc_a.item())
    assert abs(acc_b.item() - got.double().sum().item()) <= 1e-5 * abs(acc_b.item())


def test_nchw_to_nhwc_bf16_slice():
    ops = _ops()
    torch.manual_seed(5)
    x = torch.randn(2, 81, 6, 10, device=DEV)
    out = torch.zeros(2, 6, 10, 192, dtype=torch.bfloat16, device=DEV)
    ops.nchw_to_nhwc_bf16(x, out, 81)
    assert (out[..., 81:162].float() == x.permute(0, 2, 3, 1).to(torch.bfloat16).float()).all()
    assert (out[..., :81] == 0).all() and (out[..., 162:] == 0).all()


@pytest.mark.parametrize("in_ch,shape", [(3, (2, 3, 24, 40)), (1, (3, 1, 16, 16)), (3, (1, 3, 5, 7))])
def test_berk_autoencoder_tc_matches_torch_fp32(in_ch, shape):
    """SubbandAutoEncoderBerk on the 3xTF32 tensor-core chain vs the same module through torch fp32 convs (TF32 off):
    both are compared with a float64 run of the same module: the tensor-core chain must be as accurate as the fp32 one."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.layers.lifting_dwt_nets import SubbandAutoEncoderBerk
    torch.manual_seed(21)
    ae = SubbandAutoEncoderBerk(in_ch)
    with torch.no_grad():
        for prm in ae.parameters():            # non-trivial weights, GDN parameters kept near their (valid) init
            if prm.dim() == 4:
                prm.copy_(torch.randn_like(prm) * (2.0 / (prm[0].numel() ** 0.5)))
            elif prm.dim() == 2:
                prm.add_(torch.rand_like(prm) * 0.05)
    import copy
    ae64 = copy.deepcopy(ae).double().to(DEV).eval()
    ae = ae.to(DEV).eval()
    x = (torch.randn(*shape) * 2).to(DEV)
    with torch.no_grad():
        for name in ("encode", "decode"):
            ref = getattr(ae64, "ae_down" if name == "encode" else "ae_up")(x.double())     # float64 ground truth
            got = getattr(ae, name)(x)                                        # the product path (3xTF32 chain)
            t32 = ae._exact(ae.ae_down if name == "encode" else ae.ae_up, x)  # the module's torch layers, TF32 off
            scale = ref.abs().max().item()
            err_tc = (got.double() - ref).abs().max().item() / scale
            err_t32 = (t32.double() - ref).abs().max().item() / scale
            print(f"{name}: 3xTF32 chain err {err_tc:.2e}, torch fp32 err {err_t32:.2e}")
            # encode (decides the symbols): fp32-conv-chain level.  decode: the tensor core's FP32 accumulator rounds
            # toward zero, a bias that grows with the 324-648 accumulation steps of these K = 864 / 1728 convs and that
            # the inverse GDN (y * sqrt(norm)) does not normalise away like the forward GDN does; the reconstruction
            # tolerance of the path is 1e-4, the chain must stay well inside it.
            # (the small-term products go to a second accumulator, which cut this bias 3x: decode measures 0.4-1.4e-5)
            assert err_tc <= (max(3 * err_t32, 1e-5) if name == "encode" else 2.5e-5), (name, err_tc, err_t32)


@pytest.mark.parametrize("C,Cout,shape", [(96, 3, (2, 24, 40)), (32, 1, (3, 16, 16)), (96, 3, (1, 5, 7)), (32, 1, (1, 9, 33))])
def test_nhwc_split_tail_conv_matches_torch(C, Cout, shape):
    """Last conv of SubbandAutoEncoderBerk straight from the chain's [hi | lo] channels-last layout (ll_nhwc_split_conv3)
    against torch conv2d on hi + lo in float64; fp32 FMA accumulation over K = 9 C: a few 1e-7 of the output scale."""
    ops = _ops()
    torch.manual_seed(C + Cout)
    B, H, W = shape
    v = torch.randn(B, H, W, C, device=DEV)
    hi = (v.view(torch.int32) & ~0x1FFF).view(torch.float32)        # a TF32-representable high half and the rest
    z = torch.cat([hi, v - hi], dim=3).contiguous()
    w = torch.randn(Cout, C, 3, 3, device=DEV) * 0.1
    b = torch.randn(Cout, device=DEV)
    got = ops.nhwc_split_conv3(z, w, b)
    ref = F.conv2d((hi.double() + (v - hi).double()).permute(0, 3, 1, 2), w.double(), b.double(), padding=1)
    assert got.shape == (B, Cout, H, W)
    assert (got.double() - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
    nob = ops.nhwc_split_conv3(z, w, None)
    assert (nob.double() - (ref - b.double().view(1, -1, 1, 1))).abs().max().item() <= 2e-6 * ref.abs().max().item()
    with pytest.raises(Exception):
        ops.nhwc_split_conv3(z[..., :2 * C - 2].contiguous(), w, b)


@pytest.mark.parametrize("C,N,taps,epi,shape", [(96, 192, 9, 1, (2, 37, 50)), (192, 96, 9, 1, (1, 24, 40)), (96, 96, 1, 2, (3, 16, 16)),
                                                (256, 256, 9, 3, (1, 9, 33)), (32, 64, 9, 1, (1, 8, 16)), (64, 32, 1, 2, (1, 5, 7))])
def test_igemm_tf32_cta_pairs_match_single_cta_kernel(C, N, taps, epi, shape):
    """``ll_igemm_tf32`` on CTA pairs (clusters of 2, tcgen05 cta_group::2, every k-block fetched once, weight tile shared)
    against the single-CTA kernel: same MMA order per accumulator, so the outputs must be identical bit for bit -- on even
    and odd tile counts, ragged planes, 3x3 and 1x1 (GDN) instances, both accumulator plans (N <= 128 and N > 128)."""
    ops = _ops()
    torch.manual_seed(C * 7 + N + taps)
    B, H, W = shape
    v = torch.randn(B, H, W, C, device=DEV)
    hi = (v.view(torch.int32) & ~0x1FFF).view(torch.float32)
    a = torch.cat([hi, v - hi], dim=3).contiguous()
    k = 3 if taps == 9 else 1
    wt = torch.randn(N, C, k, k, device=DEV) * (1.0 / (C * taps) ** 0.5)
    wp = ops.pack_tf32_weight(wt)
    bias = torch.rand(N, device=DEV) + (1.0 if epi == 2 else 0.0)
    if epi == 2:
        with torch.no_grad():
            wp.abs_()                                   # a GDN norm: non-negative weights on squares, positive beta
            a.abs_()
    y_in = torch.randn(B, H, W, N, device=DEV) if epi == 2 else None
    y1, z1 = ops.igemm_tf32(a, wp, bias, N, epi=epi, y=y_in, pair=False)
    y2, z2 = ops.igemm_tf32(a, wp, bias, N, epi=epi, y=y_in.clone() if epi == 2 else None, pair=True)
    torch.cuda.synchronize()
    if epi != 2:
        assert torch.equal(y1, y2)
        ref = F.conv2d(v.double().permute(0, 3, 1, 2), wt.double(), bias.double(), padding=k // 2).permute(0, 2, 3, 1)
        assert (y2.double() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()      # fp32-level (K up to 2304 terms)
    if epi != 3:
        assert torch.equal(z1, z2)
        assert torch.isfinite(z2).all()


@pytest.mark.parametrize("C,N,taps,inverse,shape", [(96, 192, 9, False, (2, 37, 50)), (192, 96, 9, False, (1, 24, 40)), (96, 192, 9, True, (1, 16, 32)),
                                                    (32, 64, 9, False, (3, 16, 16)), (64, 32, 9, True, (1, 9, 33)), (96, 96, 1, False, (1, 8, 16))])
def test_fused_conv_gdn_matches_two_kernel_chain_and_float64(C, N, taps, inverse, shape):
    """``ll_igemm_tf32_gdn`` (conv + GDN in one kernel, the conv output / its square / the norm resident in tensor memory)
    against (1) the two-kernel chain ``ll_igemm_tf32`` epi 1 -> epi 2 and (2) a float64 evaluation of
    y * rsqrt(beta + gamma . y^2): fp32-level accuracy, both GDN directions, every tensor-memory plan (N = 32, 64, 96, 192)."""
    ops = _ops()
    torch.manual_seed(C + 3 * N + taps)
    B, H, W = shape
    v = torch.randn(B, H, W, C, device=DEV)
    hi = (v.view(torch.int32) & ~0x1FFF).view(torch.float32)
    a = torch.cat([hi, v - hi], dim=3).contiguous()
    k = 3 if taps == 9 else 1
    wt = torch.randn(N, C, k, k, device=DEV) * (1.0 / (C * taps) ** 0.5)
    bias = torch.randn(N, device=DEV) * 0.1
    gamma = (torch.rand(N, N, device=DEV) * 0.02 + 0.1 * torch.eye(N, device=DEV)).reshape(N, N, 1, 1).contiguous()
    beta = torch.rand(N, device=DEV) + 0.5
    wp, gp = ops.pack_tf32_weight(wt), ops.pack_tf32_weight(gamma)
    z = ops.igemm_tf32_gdn(a, wp, bias, gp, beta, N, inverse=inverse)
    y, s = ops.igemm_tf32(a, wp, bias, N, epi=1)
    _, z2 = ops.igemm_tf32(s, gp, beta, N, epi=2, inverse=inverse, y=y)
    torch.cuda.synchronize()
    got, two = z[..., :N] + z[..., N:], z2[..., :N] + z2[..., N:]
    y64 = F.conv2d(v.double().permute(0, 3, 1, 2), wt.double(), bias.double(), padding=k // 2)
    norm = F.conv2d(y64 ** 2, gamma.double(), beta.double())
    ref = (y64 * (norm.sqrt() if inverse else norm.rsqrt())).permute(0, 2, 3, 1)
    scale = ref.abs().max().item()
    e_f, e_2 = (got.double() - ref).abs().max().item() / scale, (two.double() - ref).abs().max().item() / scale
    print(f"C={C} N={N} taps={taps} inverse={inverse}: fused {e_f:.2e}, two-kernel chain {e_2:.2e} of the output scale vs float64")
    assert e_f <= max(2 * e_2, 4e-6), (e_f, e_2)
    assert torch.isfinite(z).all()
    # [hi | lo] layout: hi is TF32-representable, lo the TF32 rounding of the rest
    assert torch.equal((z[..., :N].view(torch.int32) & 0x1FFF), torch.zeros_like(z[..., :N], dtype=torch.int32))


@pytest.mark.parametrize("iC,N,inverse,shape", [(3, 96, False, (2, 37, 50)), (1, 32, False, (3, 16, 16)), (3, 96, True, (1, 9, 33)), (1, 32, True, (1, 8, 16))])
def test_fused_head_conv_gdn_matches_float64(iC, N, inverse, shape):
    """``ll_conv3_gdn_head``: first layer of SubbandAutoEncoderBerk (3x3 conv of the 1- or 3-channel subband tensor as a 3xTF32
    split on the tensor cores, windows staged in tensor memory by the epilogue warps) + its GDN, against float64."""
    ops = _ops()
    torch.manual_seed(iC + N)
    B, H, W = shape
    x = torch.randn(B, iC, H, W, device=DEV) * 2
    w0 = torch.randn(N, iC, 3, 3, device=DEV) * (1.0 / (9 * iC) ** 0.5)
    bias = torch.randn(N, device=DEV) * 0.1
    gamma = (torch.rand(N, N, device=DEV) * 0.02 + 0.1 * torch.eye(N, device=DEV)).reshape(N, N, 1, 1).contiguous()
    beta = torch.rand(N, device=DEV) + 0.5
    z = ops.conv3_gdn_head(x, w0, bias, ops.pack_tf32_weight(gamma), beta, inverse=inverse)
    torch.cuda.synchronize()
    y64 = F.conv2d(x.double(), w0.double(), bias.double(), padding=1)
    norm = F.conv2d(y64 ** 2, gamma.double(), beta.double())
    ref = (y64 * (norm.sqrt() if inverse else norm.rsqrt())).permute(0, 2, 3, 1)
    got = z[..., :N].double() + z[..., N:].double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    print(f"head iC={iC} N={N} inverse={inverse}: {err:.2e} of the output scale vs float64")
    assert err <= 3e-6, err


@pytest.mark.parametrize("key,inn,shape", [("xo", 3, (5, 32, 48)), ("xe", 1, (3, 32, 48)), ("xo", 3, (2, 5, 7)), ("xe", 1, (70, 8, 16))])
def test_coarse_causal_chain_on_the_tensor_path_matches_the_fp32_chain(key, inn, shape):
    """Coarsest-level causal chains of conditioned2ZT (masked 3x3 convs inn -> 81 inn -> 81 inn -> 27 inn -> 9 inn -> 2 inn,
    LiftingBasedDWT_net.py:298-317 of the reference) with the two dense layers as grouped 9-tap tcgen05 GEMMs on bf16
    (``_chain_bits_input``) against the exact-fp32 direct-conv chain on the same weights: the (sigma, mu) map may differ
    by bf16 operand rounding only (two layers, 2^-9 per operand), and the masks must hold (a causal chain's output at a
    pixel does not change when later pixels change)."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models import LiftingBasedDWT_net as M
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.utils import config as C
    torch.manual_seed(7 + inn)
    layer = M.DWTConditioned2EntropyLayerZTsepSubbands(C.default_config(entropy_layer="conditioned2ZTsepSubbands", dwtlevels=2)).to(DEV).eval()
    seq = layer.csc_list[-1] if key == "xo" else layer.csc_xe
    with torch.no_grad():
        for m in seq:
            if hasattr(m, "weight"):
                m.weight.mul_(3.0)          # default init is small: make every layer matter
        B, H, W = shape
        q = torch.randint(-6, 7, (B, inn, H, W), device=DEV).float()
        ref = M._chain(seq, q)
        got = layer._chain_bits_input(key, seq, q)
        assert got.shape == ref.shape == (B, 2 * inn, H, W)
        scale = ref.abs().max().item()
        err = (got - ref).abs().max().item() / scale
        print(f"chain {key} {shape}: {err:.2e} of the output scale")
        assert err <= 2e-2, err
        assert (got - ref).abs().mean().item() <= 3e-3 * scale
        # causality: changing q at and after (y0, x0) in raster order leaves the outputs before it untouched
        y0, x0 = H // 2, W // 2
        q2 = q.clone()
        q2[:, :, y0, x0:] += 5.0
        q2[:, :, y0 + 1:, :] -= 3.0
        got2 = layer._chain_bits_input(key, seq, q2)
        assert torch.equal(got2[:, :, :y0, :], got[:, :, :y0, :])
        assert torch.equal(got2[:, :, y0, :x0 + 1], got[:, :, y0, :x0 + 1])   # type-A first layer: pixel (y0, x0) itself is not an input
        M.CTX_TC_CHAIN = False
        try:
            assert torch.equal(layer._chain_bits_input(key, seq, q), ref)
        finally:
            M.CTX_TC_CHAIN = True
