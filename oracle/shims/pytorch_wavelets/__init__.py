"""pytorch_wavelets.DWTForward/DWTInverse for wave='bior4.4', mode='periodization'
(the only combination the reference uses, lifting_dwt_nets.py:228-231)."""
import torch
import torch.nn as nn

from oracle import thirdparty as tp


class DWTForward(nn.Module):
    def __init__(self, J=1, wave="db1", mode="zero"):
        super().__init__()
        if wave != "bior4.4" or mode not in ("periodization", "per"):
            raise NotImplementedError("oracle shim covers bior4.4 / periodization only")
        h0, h1, _, _ = tp.dwt97_filters()
        self.register_buffer("h0_col", h0.reshape(1, 1, -1, 1))
        self.register_buffer("h1_col", h1.reshape(1, 1, -1, 1))
        self.register_buffer("h0_row", h0.reshape(1, 1, 1, -1))
        self.register_buffer("h1_row", h1.reshape(1, 1, 1, -1))
        self.J = J
        self.mode = mode

    def forward(self, x):
        return tp.dwt97_forward(x, self.J)


class DWTInverse(nn.Module):
    def __init__(self, wave="db1", mode="zero"):
        super().__init__()
        if wave != "bior4.4" or mode not in ("periodization", "per"):
            raise NotImplementedError("oracle shim covers bior4.4 / periodization only")
        _, _, g0, g1 = tp.dwt97_filters()
        self.register_buffer("g0_col", g0.reshape(1, 1, -1, 1))
        self.register_buffer("g1_col", g1.reshape(1, 1, -1, 1))
        self.register_buffer("g0_row", g0.reshape(1, 1, 1, -1))
        self.register_buffer("g1_row", g1.reshape(1, 1, 1, -1))
        self.mode = mode

    def forward(self, coeffs):
        yl, yh = coeffs
        return tp.dwt97_inverse(yl, yh)
