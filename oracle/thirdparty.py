"""Restatement of the third-party arithmetic the reference's hot path relies on.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The packages are absent
from ``/root/reference`` and cannot be installed offline:

* ``compressai==1.2.1`` (``requirements.txt:2``) -- ``GaussianConditional``,
  ``EntropyBottleneck``, ``GDN``, ``LowerBound``, ``RGB2YCbCr``/``YCbCr2RGB``.
  Call sites: ``graphs/models/LiftingBasedDWT_net.py:3,204,209,291,307,318,
  330-365,689-690,800-832``; ``graphs/layers/lifting_dwt_nets.py:80,140-148``;
  ``agents/liftingDWT_agent.py:10,19-20,86,91``.
* ``pytorch_wavelets`` (unpinned, ``requirements.txt:36``) with
  ``PyWavelets==1.3.0`` (``requirements.txt:19``) -- ``DWTForward/DWTInverse``
  for ``wave='bior4.4', mode='periodization'`` only.  Call sites:
  ``graphs/layers/lifting_dwt_nets.py:211,230-231,250,274``.

Everything is plain torch (CPU) so that the arithmetic order matches what the
reference's own CPU path does (the reference *is* torch-on-CPU in test mode).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# compressai.ops.LowerBound (same rule as the in-repo copy utils/bound_ops.py:22-28)
# ----------------------------------------------------------------------------


class _LowerBoundFn(torch.autograd.Function):
    """max(x, bound); gradient passes where x >= bound or it pushes x upward."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        pass_through = (x >= bound) | (grad_output < 0)
        return pass_through * grad_output, None


def lower_bound(x, bound):
    if not torch.is_tensor(bound):
        bound = torch.tensor([float(bound)], dtype=x.dtype, device=x.device)
    return _LowerBoundFn.apply(x, bound)


# ----------------------------------------------------------------------------
# compressai.entropy_models.EntropyModel.quantize / GaussianConditional
# ----------------------------------------------------------------------------


def quantize(inputs, mode, means=None):
    """EntropyModel.quantize of compressai 1.2.1.

    "noise": x + U(-1/2, 1/2) drawn with ``torch.empty_like(x).uniform_``;
    "dequantize": round(x - m) + m (torch.round = half to even);
    "symbols": int(round(x - m)).
    """
    if mode == "noise":
        return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
    out = inputs.clone()
    if means is not None:
        out -= means
    out = torch.round(out)
    if mode == "dequantize":
        if means is not None:
            out += means
        return out
    if mode != "symbols":
        raise ValueError(f'Invalid quantization mode: "{mode}"')
    return out.int()


def standardized_cumulative(z):
    """Phi(z) = 0.5 * erfc(-z / sqrt(2)), as GaussianConditional writes it."""
    return 0.5 * torch.erfc(float(-(2 ** -0.5)) * z)


def gaussian_likelihood(y, scales, means=None, scale_bound=0.11):
    """GaussianConditional._likelihood: mass of N(mean, scale) on [y-1/2, y+1/2]."""
    values = y - means if means is not None else y
    scales = lower_bound(scales, scale_bound)
    values = torch.abs(values)
    upper = standardized_cumulative((0.5 - values) / scales)
    lower = standardized_cumulative((-0.5 - values) / scales)
    return upper - lower


def gaussian_conditional_forward(x, scales, means, training, scale_bound=0.11,
                                 likelihood_bound=1e-9):
    """GaussianConditional.forward(inputs, scales, means, training)."""
    y = quantize(x, "noise" if training else "dequantize", means)
    lik = gaussian_likelihood(y, scales, means, scale_bound)
    if likelihood_bound > 0:
        lik = lower_bound(lik, likelihood_bound)
    return y, lik


# ----------------------------------------------------------------------------
# compressai.entropy_models.EntropyBottleneck (factorized prior)
# ----------------------------------------------------------------------------

EB_FILTERS = (3, 3, 3, 3)
EB_INIT_SCALE = 10.0
EB_TAIL_MASS = 1e-9


def eb_init_params(channels, filters=EB_FILTERS, init_scale=EB_INIT_SCALE):
    """Parameter construction order/values of EntropyBottleneck.__init__ (1.2.1).

    Returns an ordered dict name -> tensor (matrices filled with a constant,
    biases U(-0.5,0.5) drawn from the global torch RNG in this order, factors
    zero, quantiles (-s,0,s)).
    """
    params = {}
    f = (1,) + tuple(filters) + (1,)
    scale = init_scale ** (1 / (len(filters) + 1))
    for i in range(len(filters) + 1):
        init = np.log(np.expm1(1 / scale / f[i + 1]))
        m = torch.Tensor(channels, f[i + 1], f[i])
        m.data.fill_(init)
        params[f"_matrix{i:d}"] = m
        b = torch.Tensor(channels, f[i + 1], 1)
        torch.nn.init.uniform_(b, -0.5, 0.5)
        params[f"_bias{i:d}"] = b
        if i < len(filters):
            fa = torch.Tensor(channels, f[i + 1], 1)
            torch.nn.init.zeros_(fa)
            params[f"_factor{i:d}"] = fa
    q = torch.Tensor([-init_scale, 0, init_scale]).repeat(channels, 1, 1)
    params["quantiles"] = q
    return params


def eb_target(tail_mass=EB_TAIL_MASS):
    t = np.log(2 / tail_mass - 1)
    return torch.Tensor([-t, 0, t])


def eb_logits_cumulative(p, v, n_layers=5, stop_gradient=False):
    """Per-channel monotone MLP: v (C,1,N) -> logits (C,1,N).

    for i: v = softplus(M_i) @ v + b_i ; if i < last: v += tanh(f_i) * tanh(v)
    """
    logits = v
    for i in range(n_layers):
        m = p[f"_matrix{i:d}"]
        b = p[f"_bias{i:d}"]
        if stop_gradient:
            m, b = m.detach(), b.detach()
        logits = torch.matmul(F.softplus(m), logits)
        logits = logits + b
        if i < n_layers - 1:
            fa = p[f"_factor{i:d}"]
            if stop_gradient:
                fa = fa.detach()
            logits = logits + torch.tanh(fa) * torch.tanh(logits)
    return logits


def eb_likelihood(p, y):
    lower = eb_logits_cumulative(p, y - 0.5)
    upper = eb_logits_cumulative(p, y + 0.5)
    sign = -torch.sign(lower + upper)
    sign = sign.detach()
    return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))


def entropy_bottleneck_forward(p, x, training, likelihood_bound=1e-9):
    """EntropyBottleneck.forward on (B,C,H,W): channel-major flatten, quantise
    around the per-channel median (quantiles[:,0,1]) in eval / add noise in
    training, evaluate the factorized likelihood, undo the permutation."""
    perm = [1, 0] + list(range(2, x.dim()))
    xp = x.permute(*perm).contiguous()
    shape = xp.size()
    values = xp.reshape(xp.size(0), 1, -1)
    medians = p["quantiles"][:, :, 1:2].detach()
    y = quantize(values, "noise" if training else "dequantize", medians)
    lik = eb_likelihood(p, y)
    if likelihood_bound > 0:
        lik = lower_bound(lik, likelihood_bound)
    y = y.reshape(shape).permute(*perm).contiguous()
    lik = lik.reshape(shape).permute(*perm).contiguous()
    return y, lik


def eb_loss(p):
    logits = eb_logits_cumulative(p, p["quantiles"], stop_gradient=True)
    return torch.abs(logits - eb_target().to(logits)).sum()


# ----------------------------------------------------------------------------
# compressai.layers.GDN (== in-repo graphs/layers/gdn.py:54-92, utils/parametrizers.py:31-47)
# ----------------------------------------------------------------------------

GDN_REPARAM_OFFSET = 2.0 ** -18
GDN_PEDESTAL = GDN_REPARAM_OFFSET ** 2
GDN_BETA_MIN = 1e-6


def nonneg_init(x):
    ped = torch.Tensor([GDN_PEDESTAL])
    return torch.sqrt(torch.max(x + ped, ped))


def nonneg_reparam(x, minimum=0.0):
    bound = (minimum + GDN_PEDESTAL) ** 0.5
    out = lower_bound(x, bound)
    return out ** 2 - torch.Tensor([GDN_PEDESTAL]).to(x)


def gdn(x, beta, gamma, inverse=False):
    C = x.size(1)
    b = nonneg_reparam(beta, GDN_BETA_MIN)
    g = nonneg_reparam(gamma, 0.0).reshape(C, C, 1, 1)
    norm = F.conv2d(x ** 2, g, b)
    norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
    return x * norm


# ----------------------------------------------------------------------------
# compressai.transforms colour conversion (BT.709, full range, chroma + 0.5)
# ----------------------------------------------------------------------------

KR, KG, KB = 0.2126, 0.7152, 0.0722


def rgb2ycbcr(rgb):
    r, g, b = rgb.chunk(3, -3)
    y = KR * r + KG * g + KB * b
    cb = 0.5 * (b - y) / (1 - KB) + 0.5
    cr = 0.5 * (r - y) / (1 - KR) + 0.5
    return torch.cat((y, cb, cr), dim=-3)


def ycbcr2rgb(ycbcr):
    y, cb, cr = ycbcr.chunk(3, -3)
    r = y + (2 - 2 * KR) * (cr - 0.5)
    b = y + (2 - 2 * KB) * (cb - 0.5)
    g = (y - KR * r - KB * b) / KG
    return torch.cat((r, g, b), dim=-3)


# ----------------------------------------------------------------------------
# pytorch_wavelets DWTForward / DWTInverse, wave='bior4.4', mode='periodization'
# ----------------------------------------------------------------------------

# PyWavelets bior4.4, pinned in-repo as get_cdf97_filters (lifting_dwt_nets.py:415-418).
BIOR44_DEC_LO = [0.0, 0.037828455507264, -0.023849465019557, -0.110624404418437, 0.377402855612831,
                 0.852698679008894, 0.377402855612831, -0.110624404418437, -0.023849465019557,
                 0.037828455507264]
BIOR44_DEC_HI = [0.0, -0.064538882628697, 0.040689417609164, 0.418092273221617, -0.788485616405583,
                 0.418092273221617, 0.040689417609164, -0.064538882628697, 0.0, 0.0]
BIOR44_REC_LO = [0.0, -0.064538882628697, -0.040689417609164, 0.418092273221617, 0.788485616405583,
                 0.418092273221617, -0.040689417609164, -0.064538882628697, 0.0, 0.0]
BIOR44_REC_HI = [0.0, -0.037828455507264, -0.023849465019557, 0.110624404418437, 0.377402855612831,
                 -0.852698679008894, 0.377402855612831, 0.110624404418437, -0.023849465019557,
                 -0.037828455507264]


def _analysis_1d(x, h0, h1, dim):
    """One periodised analysis stage along ``dim`` (2 or 3) of (B,C,H,W).

    pytorch_wavelets lowlevel.afb1d for mode 'per': roll by -L/2, correlate with
    the (already reversed) filters at stride 2 with L-1 zero padding, fold the
    L/2 overhanging outputs back onto the start.  Output channels are
    [lo, hi] per input channel.
    """
    C = x.shape[1]
    L = h0.numel()
    L2 = L // 2
    shape = [1, 1, 1, 1]
    shape[dim] = L
    h = torch.cat([h0.reshape(*shape), h1.reshape(*shape)] * C, dim=0)
    N = x.shape[dim]
    assert N % 2 == 0, "hot path only handles even sizes (H, W divisible by 2^L)"
    x = torch.roll(x, -L2, dims=dim)
    if dim == 2:
        lohi = F.conv2d(x, h, padding=(L - 1, 0), stride=(2, 1), groups=C)
        N2 = N // 2
        lohi[:, :, :L2] = lohi[:, :, :L2] + lohi[:, :, N2:N2 + L2]
        lohi = lohi[:, :, :N2]
    else:
        lohi = F.conv2d(x, h, padding=(0, L - 1), stride=(1, 2), groups=C)
        N2 = N // 2
        lohi[:, :, :, :L2] = lohi[:, :, :, :L2] + lohi[:, :, :, N2:N2 + L2]
        lohi = lohi[:, :, :, :N2]
    return lohi


def _synthesis_1d(lo, hi, g0, g1, dim):
    """pytorch_wavelets lowlevel.sfb1d for mode 'per'."""
    C = lo.shape[1]
    L = g0.numel()
    shape = [1, 1, 1, 1]
    shape[dim] = L
    N = 2 * lo.shape[dim]
    s = (2, 1) if dim == 2 else (1, 2)
    g0 = torch.cat([g0.reshape(*shape)] * C, dim=0)
    g1 = torch.cat([g1.reshape(*shape)] * C, dim=0)
    y = F.conv_transpose2d(lo, g0, stride=s, groups=C) + F.conv_transpose2d(hi, g1, stride=s, groups=C)
    if dim == 2:
        y[:, :, :L - 2] = y[:, :, :L - 2] + y[:, :, N:N + L - 2]
        y = y[:, :, :N]
    else:
        y[:, :, :, :L - 2] = y[:, :, :, :L - 2] + y[:, :, :, N:N + L - 2]
        y = y[:, :, :, :N]
    return torch.roll(y, 1 - L // 2, dims=dim)


def dwt97_filters(dtype=torch.float32):
    """(h0, h1) analysis -- reversed, as prep_filt_afb2d does -- and (g0, g1) synthesis."""
    h0 = torch.tensor(BIOR44_DEC_LO[::-1], dtype=dtype)
    h1 = torch.tensor(BIOR44_DEC_HI[::-1], dtype=dtype)
    g0 = torch.tensor(BIOR44_REC_LO, dtype=dtype)
    g1 = torch.tensor(BIOR44_REC_HI, dtype=dtype)
    return h0, h1, g0, g1


def dwt97_forward(x, J):
    """DWTForward(J, mode='periodization', wave='bior4.4')(x) -> (Yl, [Yh_j (B,C,3,h,w)]).

    Width axis first, then height; Yh index 0 = high along H / low along W (LH),
    1 = HL, 2 = HH; finest level first.
    """
    h0, h1, _, _ = dwt97_filters(x.dtype)
    yh = []
    ll = x
    for _ in range(J):
        lohi = _analysis_1d(ll, h0, h1, 3)
        y = _analysis_1d(lohi, h0, h1, 2)
        s = y.shape
        y = y.reshape(s[0], -1, 4, s[-2], s[-1])
        ll = y[:, :, 0].contiguous()
        yh.append(y[:, :, 1:].contiguous())
    return ll, yh


def dwt97_inverse(yl, yh):
    """DWTInverse(mode='periodization', wave='bior4.4')((Yl, Yh))."""
    _, _, g0, g1 = dwt97_filters(yl.dtype)
    ll = yl
    for h in yh[::-1]:
        lh, hl, hh = torch.unbind(h, dim=2)
        lo = _synthesis_1d(ll, lh, g0, g1, 2)
        hi = _synthesis_1d(hl, hh, g0, g1, 2)
        ll = _synthesis_1d(lo, hi, g0, g1, 3)
    return ll


def dwt97_forward_direct(x, J):
    """Same transform written as the closed-form periodic sums (numpy, float64
    accumulate of fp32 taps) -- an independent cross-check of the conv form:

    lo[n] = sum_k dec_lo[k] x[(2n+5-k) mod N],  hi likewise.
    """
    dec_lo = np.asarray(torch.tensor(BIOR44_DEC_LO, dtype=torch.float32).numpy(), dtype=np.float64)
    dec_hi = np.asarray(torch.tensor(BIOR44_DEC_HI, dtype=torch.float32).numpy(), dtype=np.float64)

    def stage(a, axis):
        N = a.shape[axis]
        n = np.arange(N // 2)
        lo = 0.0
        hi = 0.0
        for k in range(10):
            idx = (2 * n + 5 - k) % N
            t = np.take(a, idx, axis=axis)
            lo = lo + dec_lo[k] * t
            hi = hi + dec_hi[k] * t
        return lo, hi

    ll = np.asarray(x.numpy(), dtype=np.float64)
    yh = []
    for _ in range(J):
        lo_w, hi_w = stage(ll, 3)
        ll_, lh = stage(lo_w, 2)
        hl, hh = stage(hi_w, 2)
        yh.append(np.stack([lh, hl, hh], axis=2))
        ll = ll_
    return ll, yh
