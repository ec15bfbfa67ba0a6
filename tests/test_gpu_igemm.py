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
    ref = F.conv2d(_bf(x), _bf(w), b, padding=1)
    scale = ref.abs().max().item()
    for k in range(3):
        got = cat[:, 162 * k:162 * k + 81]
        assert (got - ref[:, 81 * k:81 * k + 81]).abs().max().item() <= 2e-3 * scale
        assert (cat[:, 162 * k + 81:162 * k + 162] == 7.0).all()       # csc slots untouched
    got = nh[..., 192:192 + 243].float().permute(0, 3, 1, 2)
    assert (got - ref).abs().max().item() <= 1e-2 * scale               # bf16 output rounding
    assert (nh[..., :192] == 7.0).all() and (nh[..., 192 + 243:] == 7.0).all()


@pytest.mark.parametrize("up", [True, False])
def test_ctx_head_nhwc_matches_torch(up):
    ops = _ops()
    torch.manual_seed(4)
    B, H, W = 2, 12, 20
    con = torch.round(torch.randn(B, 3, H // 2 if up else H, W // 2 if up else W) * 3).to(DEV)
    w = ((torch.rand(243, 3, 3, 3) * 2 - 1) * 0.3).to(DEV)
    b = (torch.rand(243) - 0.5).to(DEV)
    out = ops.ctx_head_nhwc(con, w, b, upsample2=up, lrelu=True)
    assert tuple(out.shape) == (B, H, W, 256) and out.dtype == torch.bfloat16
    src = con.repeat_interleave(2, 2).repeat_interleave(2, 3) if up else con
    ref = F.leaky_relu(F.conv2d(src, w, b, padding=1), 0.01)
    got = out[..., :243].float().permute(0, 3, 1, 2)
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert (out[..., 243:] == 0).all()


def test_nchw_to_nhwc_bf16_slice():
    ops = _ops()
    torch.manual_seed(5)
    x = torch.randn(2, 81, 6, 10, device=DEV)
    out = torch.zeros(2, 6, 10, 192, dtype=torch.bfloat16, device=DEV)
    ops.nchw_to_nhwc_bf16(x, out, 81)
    assert (out[..., 81:162].float() == x.permute(0, 2, 3, 1).to(torch.bfloat16).float()).all()
    assert (out[..., :81] == 0).all() and (out[..., 162:] == 0).all()
