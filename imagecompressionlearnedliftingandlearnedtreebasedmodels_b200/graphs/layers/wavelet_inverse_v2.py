"""``wavelet_inverse_v2`` -- one inverse 2-D lifting level (reference:
graphs/layers/wavelet_inverse_v2.py:5-92) on the fused sm_100a step kernel; the
interleaves of ``reconstruct_fun`` (:40-56) are strided output views of the last
update of each half, not copies."""
import torch
import torch.nn as nn

from ... import _autograd, _torch_ref, ops
from ._packing import PackCache
from .wavelet_forward_v2 import lifting_coeff, pack_steps, step_params


class wavelet_inverse_v2(nn.Module):
    def __init__(self, P, U, resnet_coeff, liftingLevel, convBlockList, cfg, nh=0, nl=0):
        super().__init__()
        self.P = P
        self.U = U
        self.lifting_level = liftingLevel
        self.resnet_coeff = resnet_coeff
        self.convBlock = convBlockList
        self.csize = cfg.clrch
        self.scale = cfg.scale
        self.nh = nh
        self.nl = nl
        self.config = cfg
        # optional key: "tc" (conv2/conv3 on tcgen05, 3xTF32 split) | "fp32" (all layers on the FP32 FMA pipe)
        self.lift_precision = cfg.get("lift_precision", ops.DEFAULT_LIFT_PRECISION) if hasattr(cfg, "get") else getattr(cfg, "lift_precision", ops.DEFAULT_LIFT_PRECISION)
        self._cache = PackCache()

    def _blobs(self):
        return pack_steps(self._cache, self.P, self.U, self.convBlock)

    def _linear(self):
        return self.P[0].linearityFlag != 1

    def one_level_lifting(self, LL, LH, HL, HH):
        """(LL, LH, HL, HH) each (B,1,h2,w2) -> x (B,1,2 h2,2 w2) (wavelet_inverse_v2.py:20-38)."""
        return self.level(LL, torch.cat((LH, HL, HH), dim=1))

    def level(self, LL, Yh):
        scale = 1 if self.scale == 1 else 0
        params = step_params(self.P, self.U, self.convBlock, self.nh, self.nl, LL)
        fast = lambda ll, yh, *ps: ops.lift_level_inv(ll, yh, self._blobs(), self.resnet_coeff, self._linear(), scale,
                                                      self.nh if scale else None, self.nl if scale else None, precision=self.lift_precision)
        ref = lambda ll, yh, *ps: _torch_ref.lift_level_inv(ll, yh, ps, self.resnet_coeff, self._linear(), scale)
        return _autograd.run(fast, ref, [LL, Yh] + params)

    def reconstruct_fun(self, up, bot):
        """Interleave along dim 2, then transpose (wavelet_inverse_v2.py:40-56)."""
        n, c, a, b = up.shape
        recon = up.new_empty(n, c, 2 * a, b)
        recon[:, :, 0::2, :] = up
        recon[:, :, 1::2, :] = bot
        return torch.transpose(recon, 2, 3)

    def lifting_inverse_row_2_stage_lifting(self, L, H):
        """(wavelet_inverse_v2.py:68-92): optional un-scale, then steps 4,3,2,1 with minus signs."""
        blobs = self._blobs()
        if self.scale == 1:
            H = H / (lifting_coeff[4] + self.nh * 0.1)
            L = L / (lifting_coeff[5] + self.nl * 0.1)
        B, C, n, m = L.shape
        Lv, Hv = L.reshape(B * C, n, m), H.reshape(B * C, n, m)
        Lo, Ho = torch.empty_like(Lv, memory_format=torch.contiguous_format), torch.empty_like(Hv, memory_format=torch.contiguous_format)
        lin = self._linear()
        ops.lift_step([(Hv, Lv, Lo)], blobs[3], -1.0, self.resnet_coeff, lin, self.lift_precision)
        ops.lift_step([(Lo, Hv, Ho)], blobs[2], -1.0, self.resnet_coeff, lin, self.lift_precision)
        ops.lift_step([(Ho, Lo, Lo)], blobs[1], -1.0, self.resnet_coeff, lin, self.lift_precision)
        ops.lift_step([(Lo, Ho, Ho)], blobs[0], -1.0, self.resnet_coeff, lin, self.lift_precision)
        return Lo.view(B, C, n, m), Ho.view(B, C, n, m)
