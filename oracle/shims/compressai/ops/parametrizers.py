"""Shim of ``compressai.ops.parametrizers`` (compressai 1.2.1) for the test oracle.

Only ``NonNegativeParametrizer`` is on the path (GDN's beta / gamma, graphs/layers/gdn.py in the reference):
stored value v -> max(v, sqrt(minimum + offset^2))^2 - offset^2, initialised with sqrt(max(x + offset^2, offset^2)).
The buffer name ``pedestal`` and the sub-module name ``lower_bound`` are part of the checkpoint layout.
"""
import torch
from torch import nn

from .bound_ops import LowerBound


class NonNegativeParametrizer(nn.Module):
    def __init__(self, minimum=0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum, self.reparam_offset = float(minimum), float(reparam_offset)
        offset_sq = self.reparam_offset * self.reparam_offset
        self.register_buffer("pedestal", torch.Tensor([offset_sq]))
        self.lower_bound = LowerBound((self.minimum + offset_sq) ** 0.5)

    def init(self, x):
        shifted = x + self.pedestal
        return torch.clamp_min(shifted, self.pedestal).sqrt()

    def forward(self, x):
        bounded = self.lower_bound(x)
        return torch.square(bounded) - self.pedestal
