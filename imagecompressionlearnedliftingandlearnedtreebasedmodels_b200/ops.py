"""Tensor-level wrappers over the C ABI (``include/ll_api.h``).

Each function checks devices/dtypes, allocates outputs/scratch with torch, and
enqueues the CUDA work on torch's current stream.  Nothing here computes on the
host and nothing falls back to PyTorch arithmetic.
"""
import ctypes

import torch

from . import _lib
from ._lib import c_voidp, check, ptr, require_device, stream_ptr

_LAUNCHES = 0   # CUDA kernels of this library enqueued so far (bench.py's gpu_launches)


def launch_count():
    return _LAUNCHES


def _count(n):
    global _LAUNCHES
    _LAUNCHES += n


LIFT_BLOB_FLOATS = 53336
AE1_BLOB_FLOATS = 2212


def _f32c(t, name):
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------- learned lifting
LIFT_FP32, LIFT_TC, LIFT_TC16 = 0, 1, 2
# conv2 / conv3 of the lifting CNNs as a 3xFP16 split on tcgen05 (same 22 significand bits per operand as 3xTF32 at half
# the tensor time; falls back to the 3xTF32 kernel inside the library when the tanh nonlinearity is switched off)
DEFAULT_LIFT_PRECISION = "tc16"


def lift_precision_code(mode):
    """``"tc16"`` (default: conv2/conv3 on tcgen05, 3xFP16 split of pre-scaled operands), ``"tc"`` (3xTF32 split; both
    have fp32-level accuracy) or ``"fp32"`` (everything on the FP32 FMA pipe) -> the ``precision`` argument of the lifting entry points (per call; nothing is process-wide)."""
    m = {"tc": LIFT_TC, "3xtf32": LIFT_TC, "tc16": LIFT_TC16, "3xfp16": LIFT_TC16, "fp32": LIFT_FP32,
         LIFT_TC: LIFT_TC, LIFT_FP32: LIFT_FP32, LIFT_TC16: LIFT_TC16}.get(mode)
    if m is None:
        raise ValueError(f"unknown lifting precision {mode!r} (\"tc\" | \"tc16\" | \"fp32\")")
    return m


def pack_lift_step(pre_w, conv):
    """``pre_w``: convBlock[k].weight (1,1,3,1); ``conv``: dict conv1..conv4 -> (weight, bias).
    Returns the device blob of one lifting step (ll_pack_lift_step)."""
    require_device(pre_w)
    lib = _lib.load()
    args = [_f32c(pre_w.detach(), "pre_w")]
    for k in ("conv1", "conv2", "conv3", "conv4"):
        w, b = conv[k]
        args += [_f32c(w.detach(), k + ".weight"), _f32c(b.detach(), k + ".bias")]
    exp = [(3,), (16 * 25,), (16,), (16 * 16 * 25,), (16,), (16 * 16 * 25,), (16,), (16 * 25,), (1,)]
    for a, e in zip(args, exp):
        if a.numel() != e[0]:
            raise ValueError(f"pack_lift_step: parameter with {a.numel()} elements, expected {e[0]} "
                             "(the CUDA path is built for clrch=1, filtersize=5, depth_scale=2)")
    blob = torch.empty(LIFT_BLOB_FLOATS, dtype=torch.float32, device=pre_w.device)
    with torch.cuda.device(pre_w.device):
        check(lib.ll_pack_lift_step(*[ptr(a) for a in args], ptr(blob), stream_ptr()))
    _count(2)
    return blob


def _blob_array(blobs):
    if len(blobs) != 4:
        raise ValueError("need the 4 packed lifting steps")
    return (c_voidp * 4)(*[ptr(b) for b in blobs])


def lift_level_fwd(x, blobs, res_weight=0.1, linear=False, scale=0, nh=None, nl=None, ll_out=None, yh_out=None,
                   precision=DEFAULT_LIFT_PRECISION):
    """x (B,1,h,w) -> (LL (B,1,h/2,w/2), Yh (B,3,h/2,w/2) = [LH,HL,HH])."""
    require_device(x)
    x = _f32c(x, "x")
    B, C, h, w = x.shape
    if C != 1:
        raise ValueError("learned lifting runs on single-channel planes (clrch == 1)")
    if h % 2 or w % 2:
        raise ValueError(f"lift_level_fwd: h, w must be even, got {h}x{w}")
    lib = _lib.load()
    ll = ll_out if ll_out is not None else torch.empty(B, 1, h // 2, w // 2, dtype=torch.float32, device=x.device)
    yh = yh_out if yh_out is not None else torch.empty(B, 3, h // 2, w // 2, dtype=torch.float32, device=x.device)
    n = lib.ll_lift_level_scratch_floats(B, h, w)
    scratch = torch.empty(max(n, 1), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.ll_lift_level_fwd(ptr(x), x.stride(0), ptr(ll), ll.stride(0), ptr(yh), yh.stride(0), ptr(scratch),
                                    B, h, w, _blob_array(blobs), float(res_weight), int(bool(linear)), int(scale),
                                    ptr(nh), ptr(nl), lift_precision_code(precision), stream_ptr()))
    _count(8 + (6 if scale else 0))
    return ll, yh


def lift_level_inv(ll, yh, blobs, res_weight=0.1, linear=False, scale=0, nh=None, nl=None, precision=DEFAULT_LIFT_PRECISION):
    """(LL (B,1,h2,w2), Yh (B,3,h2,w2)) -> x (B,1,2*h2,2*w2)."""
    require_device(ll)
    ll = _f32c(ll, "ll")
    yh = _f32c(yh, "yh")
    B, C, h2, w2 = ll.shape
    if C != 1 or tuple(yh.shape) != (B, 3, h2, w2):
        raise ValueError(f"lift_level_inv: shapes {tuple(ll.shape)} / {tuple(yh.shape)}")
    h, w = 2 * h2, 2 * w2
    lib = _lib.load()
    x = torch.empty(B, 1, h, w, dtype=torch.float32, device=ll.device)
    n = lib.ll_lift_level_scratch_floats(B, h, w)
    scratch = torch.empty(max(n, 1), dtype=torch.float32, device=ll.device)
    with torch.cuda.device(ll.device):
        check(lib.ll_lift_level_inv(ptr(ll), ll.stride(0), ptr(yh), yh.stride(0), ptr(x), x.stride(0), ptr(scratch),
                                    B, h, w, _blob_array(blobs), float(res_weight), int(bool(linear)), int(scale),
                                    ptr(nh), ptr(nl), lift_precision_code(precision), stream_ptr()))
    _count(8 + (6 if scale else 0))
    return x


# ----------------------------------------------------------------------------- CDF 9/7
def dwt97_forward(x, J):
    """DWTForward(J, 'periodization', 'bior4.4'): x (B,C,H,W) -> (Yl (B,C,h,w), [Yh_j (B,C,3,h_j,w_j)])."""
    require_device(x)
    x = _f32c(x, "x")
    B, C, H, W = x.shape
    if J < 1 or H % (1 << J) or W % (1 << J):
        raise ValueError(f"dwt97_forward: H, W must be divisible by 2^{J}, got {H}x{W}")
    lib = _lib.load()
    N = B * C
    yh = [torch.empty(B, C, 3, H >> (j + 1), W >> (j + 1), dtype=torch.float32, device=x.device) for j in range(J)]
    yl = torch.empty(B, C, H >> J, W >> J, dtype=torch.float32, device=x.device)
    scratch = torch.empty(max(lib.ll_dwt97_scratch_floats(N, H, W, J), 1), dtype=torch.float32, device=x.device)
    arr = (c_voidp * J)(*[ptr(t) for t in yh])
    with torch.cuda.device(x.device):
        check(lib.ll_dwt97_fwd(ptr(x), ptr(yl), arr, ptr(scratch), N, H, W, J, stream_ptr()))
    _count(J)
    return yl, yh


def dwt97_inverse(yl, yh):
    """DWTInverse('periodization', 'bior4.4')((Yl, Yh))."""
    require_device(yl)
    lib = _lib.load()
    yl = _f32c(yl, "yl")
    yh = [_f32c(y, "yh") for y in yh]
    J = len(yh)
    B, C, h, w = yl.shape
    H, W = h << J, w << J
    for j, y in enumerate(yh):
        if tuple(y.shape) != (B, C, 3, H >> (j + 1), W >> (j + 1)):
            raise ValueError(f"dwt97_inverse: yh[{j}] has shape {tuple(y.shape)}, expected {(B, C, 3, H >> (j + 1), W >> (j + 1))}")
    N = B * C
    x = torch.empty(B, C, H, W, dtype=torch.float32, device=yl.device)
    scratch = torch.empty(max(lib.ll_dwt97_scratch_floats(N, H, W, J), 1), dtype=torch.float32, device=yl.device)
    arr = (c_voidp * J)(*[ptr(t) for t in yh])
    with torch.cuda.device(yl.device):
        check(lib.ll_dwt97_inv(ptr(yl), arr, ptr(x), ptr(scratch), N, H, W, J, stream_ptr()))
    _count(J)
    return x


# ----------------------------------------------------------------------------- pointwise auto-encoder
def pack_ae1(layers, C, transposed):
    """``layers``: 4 (weight, bias) pairs of ae_down (Conv2d) or ae_up (ConvTranspose2d)."""
    dev = layers[0][0].device
    require_device(layers[0][0])
    lib = _lib.load()
    args = []
    for w, b in layers:
        args += [_f32c(w.detach(), "w"), _f32c(b.detach(), "b")]
    exp = [32 * C, 32 * C, 1024 * C, 32 * C, 1024 * C, 32 * C, 32 * C, C]
    for a, e in zip(args, exp):
        if a.numel() != e:
            raise ValueError(f"pack_ae1: parameter with {a.numel()} elements, expected {e}")
    blob = torch.empty(C * AE1_BLOB_FLOATS, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.ll_pack_ae1(*[ptr(a) for a in args], C, int(bool(transposed)), ptr(blob), stream_ptr()))
    _count(1)
    return blob


def ae1_apply(x, blob, want_round=False):
    """Pointwise 1->32->32->32->1 tanh MLP per channel of x (B,C,H,W); optionally also round(y)."""
    require_device(x)
    x = _f32c(x, "x")
    B, C = x.shape[0], x.shape[1]
    n = x[0, 0].numel() if B * C else 0
    lib = _lib.load()
    y = torch.empty_like(x)
    q = torch.empty_like(x) if want_round else None
    with torch.cuda.device(x.device):
        check(lib.ll_ae1_apply(ptr(x), ptr(y), ptr(q), ptr(blob), B, C, n, stream_ptr()))
    _count(1)
    return (y, q) if want_round else y


def lift_step(jobs, blob, sign, res_weight=0.1, linear=False, precision=DEFAULT_LIFT_PRECISION):
    """One fused lifting step on up to two (src, din, dout) triples of equally shaped
    3-D strided views (B, ny, nx); the 3-tap pre-filter runs along dim 1.  dout may alias din."""
    lib = _lib.load()
    if not 1 <= len(jobs) <= 2:
        raise ValueError("lift_step takes 1 or 2 jobs")
    arr = (_lib.ll_lift_job * len(jobs))()
    dev = jobs[0][0].device
    for a, (src, din, dout) in zip(arr, jobs):
        for t in (src, din, dout):
            require_device(t)
            if t.dtype != torch.float32 or t.dim() != 3 or t.shape != src.shape:
                raise ValueError("lift_step: views must be float32, 3-D and equally shaped")
        for f, t in (("src", src), ("din", din), ("dout", dout)):
            v = getattr(a, f)
            v.ptr, v.sb, v.sy, v.sx = t.data_ptr(), t.stride(0), t.stride(1), t.stride(2)
        a.nb, a.ny, a.nx = src.shape
    with torch.cuda.device(dev):
        check(lib.ll_lift_step(arr, len(jobs), ptr(blob), float(sign), float(res_weight), int(bool(linear)),
                               lift_precision_code(precision), stream_ptr()))
    _count(1)


# ----------------------------------------------------------------------------- entropy-model kernels
EB_BLOB_FLOATS = 64


def conv2d(x, weight, bias=None, groups=1, lrelu=False, upsample2=False, out=None, co_group=0, co_stride=0, co_off=0):
    """fp32 direct conv, stride 1, zero padding K//2 (ll_conv2d).  ``x`` (B,Cin,H,W) or the
    half-resolution parent when ``upsample2``; ``out`` may be a larger tensor written through the
    channel remap ``(co // co_group) * co_stride + co_off + co % co_group``."""
    require_device(x)
    x = _f32c(x, "x")
    w = _f32c(weight.detach(), "weight")
    b = _f32c(bias.detach(), "bias") if bias is not None else None
    B, Cin, Hs, Ws = x.shape
    H, W = (2 * Hs, 2 * Ws) if upsample2 else (Hs, Ws)
    Cout, cin_g, K, K2 = w.shape
    if K != K2 or cin_g * groups != Cin:
        raise ValueError(f"conv2d: weight {tuple(w.shape)} does not fit input {tuple(x.shape)} with groups={groups}")
    if out is None:
        out = torch.empty(B, Cout, H, W, dtype=torch.float32, device=x.device)
    elif out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape[2:]) != (H, W) or out.shape[0] != B:
        raise ValueError("conv2d: bad out tensor")
    lib = _lib.load()
    with torch.cuda.device(x.device):
        check(lib.ll_conv2d(ptr(x), x.stride(0) if B > 1 else Cin * Hs * Ws, ptr(w), ptr(b), ptr(out),
                            out.stride(0) if B > 1 else out[0].numel(), B, Cin, H, W, Cout, K, groups,
                            int(bool(upsample2)), int(bool(lrelu)), co_group, co_stride, co_off, stream_ptr()))
    _count(1)
    return out


def pw_mlp3(x, conv1, conv2, conv3):
    """Conv1x1(32->32), LeakyReLU, Conv1x1(32->32), LeakyReLU, Conv1x1(32->1) of a ZTBlock dependency net in one pass
    (ll_pw_mlp3).  ``x`` (B,32,H,W); ``conv*`` the three nn.Conv2d modules.  Returns (B,1,H,W)."""
    require_device(x)
    x = _f32c(x, "x")
    B, C, H, W = x.shape
    ws = [_f32c(c.weight.detach(), "weight") for c in (conv1, conv2, conv3)]
    bs = [_f32c(c.bias.detach(), "bias") if c.bias is not None else None for c in (conv1, conv2, conv3)]
    if tuple(ws[0].shape) != (C, C, 1, 1) or tuple(ws[1].shape) != (C, C, 1, 1) or tuple(ws[2].shape) != (1, C, 1, 1) \
            or bs[0] is None or bs[1] is None:
        raise ValueError("pw_mlp3: expected biased 1x1 convs C->C, C->C, C->1")
    out = torch.empty(B, 1, H, W, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().ll_pw_mlp3(ptr(x), C * H * W, ptr(ws[0]), ptr(bs[0]), ptr(ws[1]), ptr(bs[1]), ptr(ws[2]), ptr(bs[2]),
                                     ptr(out), H * W, B, C, H * W, stream_ptr()))
    _count(1)
    return out


# ----------------------------------------------------------------------------- tcgen05 context CNNs
IG_BK = 64      # input channels per k-block (one 128-byte swizzle row of bf16)


def _pad_to(n, m):
    return (n + m - 1) // m * m


def ctx_conv_nhwc(x, weight, bias, groups=1, live_taps=None, upsample2=False, lrelu=False, out=None, coff=0,
                  co_group=0, co_gstride=0, region=None):
    """Small-Cin conv (plc head / masked csc) in fp32, written channels-last in bf16 (ll_ctx_conv_nhwc).
    ``out`` (B,H,W,Ctot) bf16 or None (allocated with ``region`` channels, default Cout padded to 64);
    the ``region`` channels from ``coff`` are all written (unmapped ones as zeros)."""
    require_device(x)
    x = _f32c(x, "x")
    w = _f32c(weight.detach(), "weight")
    b = _f32c(bias.detach(), "bias") if bias is not None else None
    B, Cin, Hs, Ws = x.shape
    H, W = (2 * Hs, 2 * Ws) if upsample2 else (Hs, Ws)
    Cout, cin_g, K, K2 = w.shape
    if K != K2 or cin_g * groups != Cin:
        raise ValueError(f"ctx_conv_nhwc: weight {tuple(w.shape)} does not fit input {tuple(x.shape)} with groups={groups}")
    if region is None:
        region = _pad_to(Cout, IG_BK) if co_group <= 0 else _pad_to((Cout // co_group) * co_gstride, 8)
    if out is None:
        out = torch.empty(B, H, W, coff + region, dtype=torch.bfloat16, device=x.device)
    elif out.dtype != torch.bfloat16 or not out.is_contiguous() or tuple(out.shape[:3]) != (B, H, W):
        raise ValueError("ctx_conv_nhwc: bad out tensor")
    with torch.cuda.device(x.device):
        check(_lib.load().ll_ctx_conv_nhwc(ptr(x), ptr(w), ptr(b), ptr(out), B, Cin, H, W, Cout, K, groups,
                                           K * K if live_taps is None else live_taps, int(bool(upsample2)),
                                           int(bool(lrelu)), out.shape[3], coff, co_group, co_gstride, region, stream_ptr()))
    _count(1)
    return out


IM2COL_CH = 320   # ll_ctx_im2col: [0,128) plc head (27 x [hi|lo|hi]), 128 + 64 g: csc group g (12 x [hi|lo|hi])


def ctx_im2col(con, q):
    """Quantised parent ``con`` (B,3,H/2,W/2) and child ``q`` (B,3,H,W) -> bf16 (B,H,W,320): the split im2col rows of the
    plc head and the masked csc for their 1-tap igemm layers (ll_ctx_im2col)."""
    require_device(q)
    con, q = _f32c(con, "con"), _f32c(q, "q")
    B, C, H, W = q.shape
    if C != 3 or tuple(con.shape) != (B, 3, H // 2, W // 2) or H % 2 or W % 2:
        raise ValueError(f"ctx_im2col: child {tuple(q.shape)} / parent {tuple(con.shape)} (3 subbands, parent at half size)")
    out = torch.empty(B, H, W, IM2COL_CH, dtype=torch.bfloat16, device=q.device)
    with torch.cuda.device(q.device):
        check(_lib.load().ll_ctx_im2col(ptr(con), ptr(q), ptr(out), B, H, W, stream_ptr()))
    _count(1)
    return out


def split_bf16_weight(w2d):
    """(N, K) fp32 -> (N, 3K) fp32 holding [W_hi | W_hi | W_lo] (W_hi = bf16(W), W_lo = bf16(W - W_hi), exact in bf16):
    the weight side of the [x_hi | x_lo | x_hi] operands ``ctx_im2col`` writes."""
    hi = w2d.to(torch.bfloat16).to(torch.float32)
    lo = (w2d - hi).to(torch.bfloat16).to(torch.float32)
    return torch.cat((hi, hi, lo), dim=1)


def pack_igemm_weight(weight, npad=None, kpad=None):
    """(Co,Ci,R,S) fp32 -> bf16 [R*S][Npad][Kpad] device blob for ``igemm_conv``."""
    require_device(weight)
    w = _f32c(weight.detach(), "weight")
    Co, Ci, R, S = w.shape
    taps = R * S
    npad = npad or _pad_to(Co, 16)
    kpad = kpad or _pad_to(Ci, IG_BK)
    wp = torch.empty(taps, npad, kpad, dtype=torch.bfloat16, device=w.device)
    with torch.cuda.device(w.device):
        check(_lib.load().ll_pack_igemm_weight(ptr(w), ptr(wp), Co, Ci, taps, npad, kpad, stream_ptr()))
    _count(1)
    return wp


def nchw_to_nhwc_bf16(x, out, coff):
    """fp32 (B,C,H,W) -> channels [coff, coff+C) of the bf16 NHWC tensor ``out`` (B,H,W,Ctot)."""
    require_device(x)
    x = _f32c(x, "x")
    B, C, H, W = x.shape
    if out.dtype != torch.bfloat16 or not out.is_contiguous() or tuple(out.shape[:3]) != (B, H, W):
        raise ValueError("nchw_to_nhwc_bf16: bad out tensor")
    with torch.cuda.device(x.device):
        check(_lib.load().ll_nchw_to_nhwc_bf16(ptr(x), x.stride(0) if B > 1 else C * H * W, ptr(out), B, C, H, W,
                                               out.shape[3], coff, stream_ptr()))
    _count(1)
    return out


def igemm_conv(x_nhwc, wp, bias, cout, lrelu=False, out=None, co_group=0, co_stride=0, co_off=0,
               out_nhwc=None, nhwc_coff=0, nhwc_gstride=0, koff=None):
    """tcgen05 implicit-GEMM conv of a bf16 NHWC tensor (B,H,W,Cin_total).  ``wp``: (taps,Npad,Kpad) from
    ``pack_igemm_weight`` or a stack (G,taps,Npad,Kpad) for ``G`` channel groups; ``koff`` (G x Kpad/64
    ints): first input channel of every 64-channel k-block per group.  ``out``: fp32 NCHW
    (B, >= G*cout mapped channels, H, W) written through the channel remap of ``conv2d``; ``out_nhwc``:
    bf16 NHWC written at ``nhwc_coff + g*nhwc_gstride`` (all Npad channels per group).  ``cout`` is per group."""
    require_device(x_nhwc)
    if x_nhwc.dtype != torch.bfloat16 or not x_nhwc.is_contiguous() or x_nhwc.dim() != 4:
        raise TypeError("igemm_conv: x must be a contiguous bf16 (B,H,W,C) tensor")
    if wp.dtype != torch.bfloat16 or not wp.is_contiguous() or wp.dim() not in (3, 4):
        raise TypeError("igemm_conv: wp must come from pack_igemm_weight")
    B, H, W, ctot = x_nhwc.shape
    G = wp.shape[0] if wp.dim() == 4 else 1
    taps, npad, kpad = wp.shape[-3:]
    kb = kpad // IG_BK
    if koff is None:
        if G != 1 or kpad != ctot:
            raise ValueError(f"igemm_conv: koff is needed (groups={G}, Kpad={kpad}, input channels={ctot})")
        karr = None
    else:
        flat = [int(v) for row in koff for v in row] if isinstance(koff[0], (list, tuple)) else [int(v) for v in koff]
        if len(flat) != G * kb:
            raise ValueError(f"igemm_conv: koff has {len(flat)} entries, expected {G}x{kb}")
        karr = (ctypes.c_int * len(flat))(*flat)
    b = _f32c(bias.detach(), "bias") if bias is not None else None
    if b is not None and b.numel() != G * cout:
        raise ValueError(f"igemm_conv: bias has {b.numel()} elements, expected {G * cout}")
    if out is None and out_nhwc is None:
        out = torch.empty(B, G * cout, H, W, dtype=torch.float32, device=x_nhwc.device)
    if out is not None and (out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape[2:]) != (H, W)
                            or out.shape[0] != B):
        raise ValueError("igemm_conv: bad out tensor")
    if out_nhwc is not None and (out_nhwc.dtype != torch.bfloat16 or not out_nhwc.is_contiguous()
                                 or tuple(out_nhwc.shape[:3]) != (B, H, W)):
        raise ValueError("igemm_conv: bad out_nhwc tensor")
    with torch.cuda.device(x_nhwc.device):
        check(_lib.load().ll_igemm_conv(
            ptr(x_nhwc), ptr(wp), ptr(b), B, H, W, ctot, kpad, npad, cout, taps, G, karr, int(bool(lrelu)),
            ptr(out), (out.stride(0) if B > 1 else out[0].numel()) if out is not None else 0, co_group, co_stride, co_off,
            ptr(out_nhwc), out_nhwc.shape[3] if out_nhwc is not None else 0, nhwc_coff, nhwc_gstride, stream_ptr()))
    _count(1)
    return out if out is not None else out_nhwc


# ----------------------------------------------------------------------------- 3xTF32 chain (SubbandAutoEncoderBerk)
def pack_tf32_weight(weight, transposed=False):
    """(Co,Ci,R,S) conv weight -- or a ConvTranspose2d weight (in,out,R,S) when ``transposed`` -- -> fp32
    (taps, Npad, 2*Kpad) blob of [hi | lo] TF32 halves for ``igemm_tf32``."""
    require_device(weight)
    w = _f32c(weight.detach(), "weight")
    if transposed:
        Ci, Co, R, S = w.shape
    else:
        Co, Ci, R, S = w.shape
    taps, npad, kpad = R * S, _pad_to(Co, 16), _pad_to(Ci, 32)
    wp = torch.empty(taps, npad, 2 * kpad, dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        check(_lib.load().ll_pack_tf32_weight(ptr(w), ptr(wp), Co, Ci, taps, npad, kpad, int(bool(transposed)), stream_ptr()))
    _count(1)
    return wp


def igemm_tf32(a_split, wp, bias, cout, epi, inverse=False, y=None, pair=True):
    """3xTF32 implicit-GEMM conv of a channels-last fp32 split tensor (B,H,W,2C).  epi 1: returns (y, split(y^2));
    epi 2 (GDN): returns (y, split(y * rsqrt(acc + bias))) with ``y`` given; epi 3: returns (y, None)."""
    require_device(a_split)
    if a_split.dtype != torch.float32 or not a_split.is_contiguous() or a_split.dim() != 4:
        raise TypeError("igemm_tf32: a_split must be a contiguous fp32 (B,H,W,2C) tensor")
    B, H, W, c2 = a_split.shape
    C = c2 // 2
    taps, npad, k2 = wp.shape
    if k2 != 2 * C:
        raise ValueError(f"igemm_tf32: packed weight has K={k2 // 2}, input has {C} channels")
    b = _f32c(bias.detach(), "bias")
    if epi == 2:
        if y is None or tuple(y.shape) != (B, H, W, cout):
            raise ValueError("igemm_tf32: the GDN epilogue needs the raw conv output y (B,H,W,Cout)")
    else:
        y = torch.empty(B, H, W, cout, dtype=torch.float32, device=a_split.device)
    sz = torch.empty(B, H, W, 2 * cout, dtype=torch.float32, device=a_split.device) if epi != 3 else None
    with torch.cuda.device(a_split.device):
        check(_lib.load().ll_igemm_tf32(ptr(a_split), ptr(wp), ptr(b), B, H, W, C, npad, cout, taps, epi, int(bool(inverse)),
                                        ptr(y), ptr(sz), int(bool(pair)), stream_ptr()))
    _count(1)
    return y, sz


def igemm_tf32_gdn(a_split, wp, bias, gp, beta, cout, inverse=False):
    """Fused 3xTF32 conv + GDN / inverse GDN on CTA pairs (ll_igemm_tf32_gdn): channels-last split tensor (B,H,W,2C) ->
    split (B,H,W,2*cout) of y * rsqrt(beta + gamma . y^2); the conv output and the norm never leave the SM."""
    require_device(a_split)
    if a_split.dtype != torch.float32 or not a_split.is_contiguous() or a_split.dim() != 4:
        raise TypeError("igemm_tf32_gdn: a_split must be a contiguous fp32 (B,H,W,2C) tensor")
    B, H, W, c2 = a_split.shape
    C = c2 // 2
    taps, npad, k2 = wp.shape
    if k2 != 2 * C or npad != cout or tuple(gp.shape) != (1, cout, 2 * cout):
        raise ValueError(f"igemm_tf32_gdn: packed weights {tuple(wp.shape)} / {tuple(gp.shape)} do not fit C={C}, N={cout}")
    b = _f32c(bias.detach(), "bias")
    bt = _f32c(beta.detach(), "beta")
    sz = torch.empty(B, H, W, 2 * cout, dtype=torch.float32, device=a_split.device)
    with torch.cuda.device(a_split.device):
        check(_lib.load().ll_igemm_tf32_gdn(ptr(a_split), ptr(wp), ptr(b), ptr(gp), ptr(bt), B, H, W, C, cout, taps,
                                            int(bool(inverse)), ptr(sz), stream_ptr()))
    _count(1)
    return sz


GDN_FUSED_WIDTHS = (32, 64, 96, 192)


def conv3_gdn_head(x, w0, bias, gp, beta, inverse=False):
    """First layer of SubbandAutoEncoderBerk fused with its GDN (ll_conv3_gdn_head): x (B,iC,H,W) fp32 NCHW, w0 (N,iC,3,3)
    -> channels-last split (B,H,W,2N) of y * rsqrt(beta + gamma . y^2), y = conv3x3(x) + bias in exact FP32 FMA."""
    require_device(x)
    x = _f32c(x, "x")
    w = _f32c(w0.detach(), "w0")
    B, iC, H, W = x.shape
    N = w.shape[0]
    if tuple(w.shape) != (N, iC, 3, 3) or tuple(gp.shape) != (1, N, 2 * N):
        raise ValueError(f"conv3_gdn_head: weights {tuple(w.shape)} / {tuple(gp.shape)} do not fit the input {tuple(x.shape)}")
    b = _f32c(bias.detach(), "bias")
    bt = _f32c(beta.detach(), "beta")
    sz = torch.empty(B, H, W, 2 * N, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().ll_conv3_gdn_head(ptr(x), ptr(w), ptr(b), ptr(gp), ptr(bt), B, iC, H, W, N, int(bool(inverse)), ptr(sz),
                                            stream_ptr()))
    _count(1)
    return sz


def nchw_to_nhwc_split(x, squares=True, want_y=True):
    """fp32 NCHW -> (y NHWC raw | None, split NHWC (B,H,W,2C) of x^2 (``squares``) or x)."""
    require_device(x)
    x = _f32c(x, "x")
    B, C, H, W = x.shape
    y = torch.empty(B, H, W, C, dtype=torch.float32, device=x.device) if want_y else None
    sz = torch.empty(B, H, W, 2 * C, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().ll_nchw_to_nhwc_split(ptr(x), ptr(y), ptr(sz), B, C, H, W, 0 if squares else 1, stream_ptr()))
    _count(1)
    return y, sz


def nhwc_split_to_nchw(z):
    """split NHWC (B,H,W,2C) -> fp32 NCHW (B,C,H,W) = hi + lo."""
    require_device(z)
    B, H, W, c2 = z.shape
    out = torch.empty(B, c2 // 2, H, W, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        check(_lib.load().ll_nhwc_split_to_nchw(ptr(z), ptr(out), B, c2 // 2, H, W, stream_ptr()))
    _count(1)
    return out


def nhwc_lrelu_conv1(y, w, bias, cin, out=None, lrelu=True):
    """Raw channels-last conv output y (B,H,W,Cpad) -> 1x1 conv of lrelu(y[..., :cin]) -> fp32 NCHW (B,Cout,H,W)."""
    require_device(y)
    if y.dtype != torch.float32 or not y.is_contiguous() or y.dim() != 4:
        raise TypeError("nhwc_lrelu_conv1: y must be a contiguous fp32 (B,H,W,Cpad) tensor")
    B, H, W, cpad = y.shape
    w = _f32c(w.detach().reshape(w.shape[0], -1), "w")
    if w.shape[1] != cin or cin > cpad:
        raise ValueError(f"nhwc_lrelu_conv1: weight {tuple(w.shape)} does not match cin={cin} (Cpad={cpad})")
    b = _f32c(bias.detach(), "bias") if bias is not None else None
    if out is None:
        out = torch.empty(B, w.shape[0], H, W, dtype=torch.float32, device=y.device)
    elif tuple(out.shape) != (B, w.shape[0], H, W) or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError("nhwc_lrelu_conv1: out must be a contiguous fp32 (B,Cout,H,W) tensor")
    with torch.cuda.device(y.device):
        check(_lib.load().ll_nhwc_lrelu_conv1(ptr(y), ptr(w), ptr(b), ptr(out), B, H * W, cin, cpad, w.shape[0], int(bool(lrelu)), stream_ptr()))
    _count(1)
    return out


RANS_GAUSS, RANS_EB, RANS_GAUSS_GRID = 0, 1, 2


def rans_streams_per_image(n_per_image, target=8192, batch=1, min_target=2048, fill_warps=148 * 8):
    """Streams per image: ~``target`` samples per stream.  A stream costs 8 bytes on top of its payload (32-bit flush +
    32-bit length entry), i.e. 0.008 bit per sample at the default -- small against the ~0.02 bit a near-certain symbol
    costs at low rates.  One warp codes one stream, so when ``batch`` images at ``target`` would leave SMs idle (fewer than
    ``fill_warps`` streams in the launch) the streams are shortened, down to ``min_target`` samples (0.03 bit per sample)."""
    n = int(n_per_image)
    s = (n + target - 1) // target
    if batch * s < fill_warps:
        s = max(s, min((n + min_target - 1) // min_target, (fill_warps + batch - 1) // max(int(batch), 1)))
    return int(max(1, min(4096, s)))


def rans_encode(mode, y, par, streams=None):
    """Entropy-code the dequantised tensor ``y`` (B,C,H,W) under the model's own distribution (``mode`` RANS_GAUSS:
    ``par`` = ms (B,2C,H,W); RANS_EB: ``par`` = ll_pack_eb blob).  Returns (words uint16 (n,), counts int32 (B*S,), S):
    stream b*S+s of image b occupies ``counts[b*S+s]`` consecutive words."""
    require_device(y)
    y = _f32c(y, "y")
    par = _f32c(par, "par")
    B, C, H, W = y.shape
    hw = H * W
    S = int(streams) if streams else rans_streams_per_image(C * hw, batch=B)
    lib = _lib.load()
    cap = int(lib.ll_rans_stream_cap(C * hw, S))
    scratch = torch.empty(B * S * cap, dtype=torch.int16, device=y.device)
    counts = torch.zeros(B * S, dtype=torch.int32, device=y.device)
    with torch.cuda.device(y.device):
        check(lib.ll_rans_encode(int(mode), ptr(y), ptr(par), B, C, hw, S, ptr(scratch), ptr(counts), stream_ptr()))
        offsets = torch.cumsum(counts.to(torch.int64), 0) - counts
        total = int(offsets[-1] + counts[-1]) if B * S else 0
        words = torch.empty(total, dtype=torch.int16, device=y.device)
        check(lib.ll_rans_pack(ptr(scratch), ptr(counts), ptr(offsets), B * S, cap, ptr(words), stream_ptr()))
    _count(2)
    return words, counts, S


def rans_decode(mode, words, counts, par, shape, S):
    """Inverse of :func:`rans_encode`: returns the dequantised tensor of ``shape`` (B,C,H,W) bit for bit."""
    require_device(words)
    par = _f32c(par, "par")
    B, C, H, W = shape
    if counts.numel() != B * S:
        raise ValueError(f"rans_decode: {counts.numel()} stream lengths for {B} images x {S} streams")
    counts = counts.to(device=words.device, dtype=torch.int32)
    offsets = (torch.cumsum(counts.to(torch.int64), 0) - counts).contiguous()
    y = torch.empty(B, C, H, W, dtype=torch.float32, device=words.device)
    with torch.cuda.device(words.device):
        check(_lib.load().ll_rans_decode(int(mode), ptr(words), ptr(offsets), ptr(par), B, C, H * W, int(S), ptr(y), stream_ptr()))
    _count(1)
    return y


def rgb_to_ycbcr_shift(rgb):
    """Agent pre-processing (agents/liftingDWT_agent.py:170-171): RGB (B,3,H,W) in [0,1] -> YCbCr (BT.709) with Y - 0.5."""
    require_device(rgb)
    rgb = _f32c(rgb, "rgb")
    if rgb.dim() != 4 or rgb.shape[1] != 3:
        raise ValueError(f"rgb_to_ycbcr_shift: expected (B,3,H,W), got {tuple(rgb.shape)}")
    out = torch.empty_like(rgb)
    with torch.cuda.device(rgb.device):
        check(_lib.load().ll_rgb_to_ycbcr_shift(ptr(rgb), ptr(out), rgb.shape[0], rgb.shape[2] * rgb.shape[3], stream_ptr()))
    _count(1)
    return out


def ycbcr_to_rgb_sse(ycc_hat, rgb_ref=None, want_xhat=True):
    """Agent post-processing (:174-186): xhat = clamp(YCbCr2RGB(yhat + (0.5,0,0)) - 0.5, +-0.5) and, with ``rgb_ref``,
    the per-image sum of squared errors against ``rgb_ref - 0.5`` (float64 (B,)).  Returns (xhat | None, sse | None)."""
    require_device(ycc_hat)
    y = _f32c(ycc_hat, "ycc_hat")
    if y.dim() != 4 or y.shape[1] != 3:
        raise ValueError(f"ycbcr_to_rgb_sse: expected (B,3,H,W), got {tuple(y.shape)}")
    ref = _f32c(rgb_ref, "rgb_ref") if rgb_ref is not None else None
    if ref is not None and ref.shape != y.shape:
        raise ValueError("ycbcr_to_rgb_sse: reference and reconstruction shapes differ")
    if ref is None and not want_xhat:
        raise ValueError("ycbcr_to_rgb_sse: nothing to compute")
    xhat = torch.empty_like(y) if want_xhat else None
    sse = torch.zeros(y.shape[0], dtype=torch.float64, device=y.device) if ref is not None else None
    with torch.cuda.device(y.device):
        check(_lib.load().ll_ycbcr_to_rgb_sse(ptr(y), ptr(ref), ptr(xhat), y.shape[0], y.shape[2] * y.shape[3], ptr(sse), stream_ptr()))
    _count(1)
    return xhat, sse


def nhwc_split_conv3(z, w, bias):
    """split NHWC (B,H,W,2C) -> 3x3 conv (zero padding 1, cross-correlation, exact fp32) -> fp32 NCHW (B,Cout,H,W)."""
    require_device(z)
    B, H, W, c2 = z.shape
    w = _f32c(w, "w")
    cout = w.shape[0]
    if w.shape[1] * 2 != c2 or tuple(w.shape[2:]) != (3, 3):
        raise ValueError(f"nhwc_split_conv3: weight {tuple(w.shape)} does not match {c2 // 2} input channels, 3x3")
    b = _f32c(bias, "bias") if bias is not None else None
    out = torch.empty(B, cout, H, W, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        check(_lib.load().ll_nhwc_split_conv3(ptr(z), ptr(w), ptr(b), ptr(out), B, c2 // 2, cout, H, W, stream_ptr()))
    _count(1)
    return out


def cgp_tail_rate(h2, w3, b3, w4, b4, x, noise=None, want_y=False, want_ms=False, acc=None):
    """Last two grouped 1x1 layers of the cgp MLP + Gaussian rate (ll_cgp_tail_rate).  h2 (B,G*C2,H,W)
    fp32; w3 (G*C3,C2,1,1); w4 (2G,C3,1,1); x (B,G,H,W).  Returns bits (+ y, + ms (B,2G,H,W))."""
    require_device(x)
    x = _f32c(x, "x")
    h2 = _f32c(h2, "h2")
    B, G, H, W = x.shape
    C2, C3 = w3.shape[1], w4.shape[1]
    if h2.shape[1] != G * C2 or w3.shape[0] != G * C3 or w4.shape[0] != 2 * G or tuple(h2.shape[2:]) != (H, W):
        raise ValueError("cgp_tail_rate: shape mismatch")
    if noise is not None:
        noise = _f32c(noise, "noise")
    w3c, b3c, w4c, b4c = (_f32c(t.detach(), "w") for t in (w3, b3, w4, b4))
    bits = torch.empty_like(x)
    y = torch.empty_like(x) if want_y else None
    ms = torch.empty(B, 2 * G, H, W, dtype=torch.float32, device=x.device) if want_ms else None
    hw = H * W
    with torch.cuda.device(x.device):
        check(_lib.load().ll_cgp_tail_rate(ptr(h2), h2.stride(0) if B > 1 else G * C2 * hw, ptr(w3c), ptr(b3c), ptr(w4c),
                                           ptr(b4c), ptr(x), G * hw, ptr(noise), ptr(bits), G * hw, ptr(y), ptr(ms),
                                           B, G, C2, C3, hw, ptr(acc), stream_ptr()))
    _count(1)
    res = (bits,) + ((y,) if want_y else ()) + ((ms,) if want_ms else ())
    return res[0] if len(res) == 1 else res


def igemm_cgp_tail(h1, wp2, bias2, c2, koff, w3, b3, w4, b4, x, noise=None, acc=None):
    """cgp layers 2-4 + Gaussian rate in one launch (ll_igemm_cgp_tail): ``h1`` bf16 (B,H,W,C) -> grouped 1x1 tcgen05 GEMM
    (``wp2`` (G,1,64,Kpad)) whose accumulator feeds layers 3-4 and the rate of ``x`` (B,G,H,W).  Returns bits (B,G,H,W);
    same arithmetic as ``igemm_conv`` + ``cgp_tail_rate`` without the fp32 map in between."""
    require_device(x)
    x = _f32c(x, "x")
    B, G, H, W = x.shape
    if h1.dtype != torch.bfloat16 or not h1.is_contiguous() or tuple(h1.shape[:3]) != (B, H, W):
        raise TypeError("igemm_cgp_tail: h1 must be a contiguous bf16 (B,H,W,C) tensor")
    if wp2.dtype != torch.bfloat16 or not wp2.is_contiguous() or wp2.dim() != 4 or wp2.shape[0] != G or wp2.shape[1] != 1 or wp2.shape[2] != 64:
        raise TypeError("igemm_cgp_tail: wp2 must be a (G,1,64,Kpad) stack from pack_igemm_weight")
    kpad = wp2.shape[3]
    kb = kpad // IG_BK
    flat = [int(v) for row in koff for v in row]
    if len(flat) != G * kb:
        raise ValueError(f"igemm_cgp_tail: koff has {len(flat)} entries, expected {G}x{kb}")
    karr = (ctypes.c_int * len(flat))(*flat)
    c3 = w4.shape[1]
    if w3.shape[0] != G * c3 or w3.shape[1] != c2 or w4.shape[0] != 2 * G:
        raise ValueError("igemm_cgp_tail: shape mismatch")
    if noise is not None:
        noise = _f32c(noise, "noise")
    b2c, w3c, b3c, w4c, b4c = (_f32c(t.detach(), "w") for t in (bias2, w3, b3, w4, b4))
    bits = torch.empty_like(x)
    hw = H * W
    with torch.cuda.device(x.device):
        check(_lib.load().ll_igemm_cgp_tail(ptr(h1), ptr(wp2), ptr(b2c), B, H, W, h1.shape[3], kpad, c2, G, karr, ptr(w3c), ptr(b3c),
                                            ptr(w4c), ptr(b4c), c3, ptr(x), G * hw, ptr(noise), ptr(bits), G * hw, ptr(acc),
                                            stream_ptr()))
    _count(1)
    return bits


def quantize(x, noise=None):
    """round-half-even(x), or x + noise in training (ll_quantize)."""
    require_device(x)
    x = _f32c(x, "x")
    if noise is not None:
        noise = _f32c(noise, "noise")
    q = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load().ll_quantize(ptr(x), ptr(noise), ptr(q), x.numel(), stream_ptr()))
    _count(1)
    return q


def new_bit_accumulator(device):
    return torch.zeros(1, dtype=torch.float64, device=device)


def gauss_rate(x, ms, noise=None, want_y=False, acc=None):
    """bits = -log2 GaussianConditional likelihood.  x (B,C,H,W); ms (B,2C,H,W) with sigma on even
    and mu on odd channels.  Returns bits (and y if ``want_y``); ``acc`` (float64[1]) += sum(bits)."""
    require_device(x)
    x = _f32c(x, "x")
    ms = _f32c(ms, "ms")
    B, C, H, W = x.shape
    if tuple(ms.shape) != (B, 2 * C, H, W):
        raise ValueError(f"gauss_rate: ms {tuple(ms.shape)} vs x {tuple(x.shape)}")
    if noise is not None:
        noise = _f32c(noise, "noise")
    bits = torch.empty_like(x)
    y = torch.empty_like(x) if want_y else None
    hw = H * W
    with torch.cuda.device(x.device):
        check(_lib.load().ll_gauss_rate(ptr(x), C * hw, ptr(ms), 2 * C * hw, ptr(noise), ptr(bits), C * hw, ptr(y),
                                        B, C, hw, ptr(acc), stream_ptr()))
    _count(1)
    return (bits, y) if want_y else bits


def pack_eb(params, C):
    """params: the 15 EntropyBottleneck tensors in registration order (see ll_pack_eb)."""
    dev = params[0].device
    require_device(params[0])
    ps = [_f32c(p.detach(), "eb param") for p in params]
    if len(ps) != 15:
        raise ValueError("pack_eb: need 15 parameter tensors (filters (3,3,3,3))")
    blob = torch.empty(C * EB_BLOB_FLOATS, dtype=torch.float32, device=dev)
    arr = (c_voidp * 15)(*[ptr(p) for p in ps])
    with torch.cuda.device(dev):
        check(_lib.load().ll_pack_eb(arr, C, ptr(blob), stream_ptr()))
    _count(1)
    return blob


def eb_rate(x, blob, noise=None, acc=None):
    """EntropyBottleneck forward: returns (y, bits) for x (B,C,H,W)."""
    require_device(x)
    x = _f32c(x, "x")
    if noise is not None:
        noise = _f32c(noise, "noise")
    B, C, H, W = x.shape
    y = torch.empty_like(x)
    bits = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load().ll_eb_rate(ptr(x), ptr(noise), ptr(blob), ptr(y), ptr(bits), B, C, H * W, ptr(acc), stream_ptr()))
    _count(1)
    return y, bits


# ----------------------------------------------------------------------------- measurement helpers (libll_probe.so)
def fma_peak_tflops(device=None, iters=4096):
    """Measured FP32 FMA-pipe peak (TFLOP/s) of the current device: register-only FFMA2 loop."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load_probe()
    out = torch.zeros(4, dtype=torch.float32, device=device)
    blocks = _lib.load().ll_sm_count() * 8
    best = 0.0
    with torch.cuda.device(device):
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check_probe(lib.ll_fma_peak_probe(ptr(out), blocks, iters, stream_ptr()))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = max(best, blocks * 256.0 * iters * 256.0 / (ms * 1e-3) / 1e12)
    return best


def tf32_peak_tflops(device=None, n=256, kblocks=16, iters=2000):
    """Measured dense TF32 tensor-pipe peak (TFLOP/s): every SM issues ``iters`` chains of ``4 * kblocks``
    tcgen05.mma kind::tf32 M128 x N x K8 on operands resident in shared memory (no global traffic)."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load_probe()
    out = torch.zeros(4, dtype=torch.float32, device=device)
    blocks = _lib.load().ll_sm_count()
    best = 0.0
    with torch.cuda.device(device):
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check_probe(lib.ll_tf32_peak_probe(ptr(out), blocks, iters, kblocks, n, stream_ptr()))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = max(best, blocks * float(iters) * kblocks * 4 * 2.0 * 128 * n * 8 / (ms * 1e-3) / 1e12)
    return best
