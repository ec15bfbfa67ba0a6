"""Module-shaped stand-ins for the third-party classes the reference builds its models
from (compressai 1.2.1 ``GDN`` / ``LowerBound`` / ``NonNegativeParametrizer`` /
``EntropyBottleneck`` / ``GaussianConditional`` / colour transforms, pytorch_wavelets
``DWTForward`` / ``DWTInverse``).  They register the same parameters and buffers under the
same names, so reference checkpoints load with ``strict=True``; their arithmetic runs on
the GPU (CUDA kernels of this package where the hot path goes through them, torch CUDA ops
for the adjacent pieces SURVEY.md section 8f lists as "next").
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _autograd, _torch_ref, ops

# --------------------------------------------------------------------------- bounds / GDN


class _LowerBoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (g < 0)) * g, None


class LowerBound(nn.Module):
    """max(x, bound) with compressai's pass-through gradient (utils/bound_ops.py:22-65)."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFn.apply(x, self.bound)


class NonNegativeParametrizer(nn.Module):
    """(utils/parametrizers.py:23-47)."""

    def __init__(self, minimum=0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        self.lower_bound = LowerBound((self.minimum + self.reparam_offset ** 2) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        out = self.lower_bound(x)
        return out ** 2 - self.pedestal


class GDN(nn.Module):
    """Generalised divisive normalisation (graphs/layers/gdn.py:41-92)."""

    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def forward(self, x):
        C = x.size(1)
        beta = self.beta_reparam(self.beta)
        gamma = self.gamma_reparam(self.gamma).reshape(C, C, 1, 1)
        norm = F.conv2d(x ** 2, gamma, beta)
        norm = torch.sqrt(norm) if self.inverse else torch.rsqrt(norm)
        return x * norm


# --------------------------------------------------------------------------- colour transforms

KR, KG, KB = 0.2126, 0.7152, 0.0722


class RGB2YCbCr:
    def __call__(self, rgb):
        r, g, b = rgb.chunk(3, -3)
        y = KR * r + KG * g + KB * b
        cb = 0.5 * (b - y) / (1 - KB) + 0.5
        cr = 0.5 * (r - y) / (1 - KR) + 0.5
        return torch.cat((y, cb, cr), dim=-3)


class YCbCr2RGB:
    def __call__(self, ycbcr):
        y, cb, cr = ycbcr.chunk(3, -3)
        r = y + (2 - 2 * KR) * (cr - 0.5)
        b = y + (2 - 2 * KB) * (cb - 0.5)
        g = (y - KR * r - KB * b) / KG
        return torch.cat((r, g, b), dim=-3)


# --------------------------------------------------------------------------- CDF 9/7 filter bank

BIOR44 = dict(
    dec_lo=[0.0, 0.037828455507264, -0.023849465019557, -0.110624404418437, 0.377402855612831, 0.852698679008894,
            0.377402855612831, -0.110624404418437, -0.023849465019557, 0.037828455507264],
    dec_hi=[0.0, -0.064538882628697, 0.040689417609164, 0.418092273221617, -0.788485616405583, 0.418092273221617,
            0.040689417609164, -0.064538882628697, 0.0, 0.0],
    rec_lo=[0.0, -0.064538882628697, -0.040689417609164, 0.418092273221617, 0.788485616405583, 0.418092273221617,
            -0.040689417609164, -0.064538882628697, 0.0, 0.0],
    rec_hi=[0.0, -0.037828455507264, -0.023849465019557, 0.110624404418437, 0.377402855612831, -0.852698679008894,
            0.377402855612831, 0.110624404418437, -0.023849465019557, -0.037828455507264])


class DWTForward(nn.Module):
    """pytorch_wavelets.DWTForward for wave='bior4.4', mode='periodization' on the sm_100a
    filter-bank kernel.  The filter buffers exist for checkpoint compatibility; the kernel
    carries the same fp32 taps as immediates (csrc/dwt97_body.cuh)."""

    def __init__(self, J=1, wave="db1", mode="zero"):
        super().__init__()
        if wave != "bior4.4" or mode not in ("periodization", "per"):
            raise NotImplementedError("only wave='bior4.4', mode='periodization' (the reference's use)")
        h0 = torch.tensor(BIOR44["dec_lo"][::-1], dtype=torch.float)
        h1 = torch.tensor(BIOR44["dec_hi"][::-1], dtype=torch.float)
        self.register_buffer("h0_col", h0.reshape(1, 1, -1, 1))
        self.register_buffer("h1_col", h1.reshape(1, 1, -1, 1))
        self.register_buffer("h0_row", h0.reshape(1, 1, 1, -1))
        self.register_buffer("h1_row", h1.reshape(1, 1, 1, -1))
        self.J = J
        self.mode = mode

    def forward(self, x):
        if _autograd.needs_grad([x]):
            def ref(x):
                yh, cur = [], x
                for _ in range(self.J):
                    cur, y = _torch_ref.dwt97_fwd_level(cur)
                    yh.append(y)
                return (cur, *yh)
            out = _autograd.run(lambda x: (lambda r: (r[0], *r[1]))(ops.dwt97_forward(x, self.J)), ref, [x])
            return out[0], list(out[1:])
        return ops.dwt97_forward(x, self.J)


class DWTInverse(nn.Module):
    def __init__(self, wave="db1", mode="zero"):
        super().__init__()
        if wave != "bior4.4" or mode not in ("periodization", "per"):
            raise NotImplementedError("only wave='bior4.4', mode='periodization' (the reference's use)")
        g0 = torch.tensor(BIOR44["rec_lo"], dtype=torch.float)
        g1 = torch.tensor(BIOR44["rec_hi"], dtype=torch.float)
        self.register_buffer("g0_col", g0.reshape(1, 1, -1, 1))
        self.register_buffer("g1_col", g1.reshape(1, 1, -1, 1))
        self.register_buffer("g0_row", g0.reshape(1, 1, 1, -1))
        self.register_buffer("g1_row", g1.reshape(1, 1, 1, -1))
        self.mode = mode

    def forward(self, coeffs):
        yl, yh = coeffs
        if _autograd.needs_grad([yl] + list(yh)):
            def ref(yl, *yh):
                cur = yl
                for y in yh[::-1]:
                    cur = _torch_ref.dwt97_inv_level(cur, y)
                return cur
            return _autograd.run(lambda yl, *yh: ops.dwt97_inverse(yl, list(yh)), ref, [yl] + list(yh))
        return ops.dwt97_inverse(yl, yh)


# --------------------------------------------------------------------------- entropy models


def draw_noise(like):
    """The reference's noise draw: ``torch.empty_like(x).uniform_(-0.5, 0.5)`` (compressai
    EntropyModel.quantize), one draw per call, in call order."""
    return torch.empty_like(like).uniform_(-0.5, 0.5)


class EntropyModel(nn.Module):
    """compressai 1.2.1 ``EntropyModel`` surface used by the reference: ``quantize`` and the
    (empty) CDF buffers that appear in checkpoints."""

    def __init__(self, likelihood_bound=1e-9, entropy_coder=None, entropy_coder_precision=16):
        super().__init__()
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    def quantize(self, inputs, mode, means=None):
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            noise = draw_noise(inputs)
            if _autograd.needs_grad([inputs]):
                return inputs + noise          # identity gradient, as compressai's additive-noise quantiser
            return ops.quantize(inputs, noise)
        if means is not None:
            q = ops.quantize(inputs - means)
            return q + means if mode == "dequantize" else q.int()
        q = ops.quantize(inputs)
        return q if mode == "dequantize" else q.int()


class GaussianConditional(EntropyModel):
    """compressai 1.2.1 ``GaussianConditional(scale_table=None, scale_bound=0.11)``; ``forward``
    returns (outputs, likelihood) like the original, ``bits`` is the fused fast path."""

    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if scale_table is not None:
            raise NotImplementedError("scale tables are only used by the serial coder (out of scope)")
        if abs(float(scale_bound) - 0.11) > 1e-12:
            raise NotImplementedError("the CUDA rate kernel is built for scale_bound = 0.11 (the reference's value)")
        self.tail_mass = float(tail_mass)
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    def bits(self, inputs, ms, training=None, want_y=False, acc=None):
        """-log2 likelihood with sigma/mu interleaved on the channels of ``ms`` (B,2C,H,W)."""
        if training is None:
            training = self.training
        noise = draw_noise(inputs) if training else None
        if _autograd.needs_grad([inputs, ms]):
            fast = lambda x, ms: ops.gauss_rate(x, ms, noise, want_y, acc)
            ref = (lambda x, ms: _torch_ref.gauss_bits(x, ms, noise)) if want_y else (lambda x, ms: _torch_ref.gauss_bits(x, ms, noise)[0])
            return _autograd.run(fast, ref, [inputs, ms])
        return ops.gauss_rate(inputs, ms, noise, want_y, acc)

    def forward(self, inputs, scales, means=None, training=None):
        if means is None:
            means = torch.zeros_like(inputs)
        ms = torch.stack((scales, means), dim=2).flatten(1, 2)   # channel 2c = sigma, 2c+1 = mu
        bits, y = self.bits(inputs, ms, training, want_y=True)
        return y, torch.exp2(-bits)


class EntropyBottleneck(EntropyModel):
    """compressai 1.2.1 ``EntropyBottleneck`` (filters (3,3,3,3), init_scale 10)."""

    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        if self.filters != (3, 3, 3, 3):
            raise NotImplementedError("the CUDA kernel is built for filters (3,3,3,3) (the reference's default)")
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        channels = self.channels
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))
        from .graphs.layers._packing import PackCache
        self._cache = PackCache()

    def _param_list(self):
        ps = []
        for i in range(5):
            ps.append(getattr(self, f"_matrix{i:d}"))
            ps.append(getattr(self, f"_bias{i:d}"))
            if i < 4:
                ps.append(getattr(self, f"_factor{i:d}"))
        ps.append(self.quantiles)
        return ps

    def _blob(self):
        ps = self._param_list()
        return self._cache.get(ps, lambda: ops.pack_eb(ps, self.channels))

    def _get_medians(self):
        return self.quantiles[:, :, 1:2].detach()

    def _logits_cumulative(self, inputs, stop_gradient):
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                matrix, bias = matrix.detach(), bias.detach()
            logits = torch.matmul(F.softplus(matrix), logits) + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def loss(self):
        """Auxiliary quantile loss (tiny, host-side bookkeeping: plain torch ops)."""
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def rate(self, x, training=None, acc=None):
        """(y, bits) through the fused CUDA kernel."""
        if training is None:
            training = self.training
        noise = None
        if training:
            # compressai draws the noise on the channel-major (C, 1, B*H*W) view of x: same stream, that layout
            B, C, H, W = x.shape
            noise = draw_noise(x.new_empty(C, 1, B * H * W)).reshape(C, B, H, W).permute(1, 0, 2, 3).contiguous()
        ps = self._param_list()
        if _autograd.needs_grad([x] + ps):
            return _autograd.run(lambda x, *p: ops.eb_rate(x, self._blob(), noise, acc),
                                 lambda x, *p: _torch_ref.eb_bits(x, p, noise), [x] + ps)
        return ops.eb_rate(x, self._blob(), noise, acc)

    def forward(self, x, training=None):
        y, bits = self.rate(x, training)
        return y, torch.exp2(-bits)
