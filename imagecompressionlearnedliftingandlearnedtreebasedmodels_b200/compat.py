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

from . import ops

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
        return ops.dwt97_inverse(yl, yh)
