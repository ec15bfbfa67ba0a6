"""``wavelet_forward_v2`` -- one forward 2-D lifting level (reference:
graphs/layers/wavelet_forward_v2.py:11-81), on the fused sm_100a step kernel.

The reference slices even/odd rows, transposes, and runs 12 P-block calls through
cuDNN; here a level is 8 launches of ``ll_lift_step`` over strided views (no split,
transpose or interleave copies), composed in C++ by ``ll_lift_level_fwd``.
"""
import torch
import torch.nn as nn

from ... import _autograd, _torch_ref, ops
from ._packing import PackCache

lifting_coeff = [-1.586134342059924, -0.052980118572961, 0.882911075530934, 0.443506852043971, 0.869864451624781,
                 1.149604398860241]  # bior4.4


def step_sources(P, U, convBlock):
    """step k -> (pre-filter conv, CNN block): 1:P[0] 2:U[0] 3:P[1] 4:U[1] (wavelet_forward_v2.py:60-74)."""
    return [(convBlock[0], P[0]), (convBlock[1], U[0]), (convBlock[2], P[1]), (convBlock[3], U[1])]


def step_params(P, U, convBlock, nh, nl, like):
    """The 38 tensors ``_torch_ref.lift_level_*`` takes: per step (pre-filter, w1, b1, ..., w4, b4), then nh, nl."""
    ts = []
    for pre, blk in step_sources(P, U, convBlock):
        ts.append(pre.weight)
        for k in ("conv1", "conv2", "conv3", "conv4"):
            ts += [getattr(blk, k).weight, getattr(blk, k).bias]
    for n in (nh, nl):
        ts.append(n if torch.is_tensor(n) else torch.zeros(1, 1, 1, 1, device=like.device, dtype=like.dtype))
    return ts


def pack_steps(cache, P, U, convBlock):
    src = step_sources(P, U, convBlock)
    tensors = []
    for pre, blk in src:
        tensors.append(pre.weight)
        for k in ("conv1", "conv2", "conv3", "conv4"):
            tensors += [getattr(blk, k).weight, getattr(blk, k).bias]
    return cache.get(tensors, lambda: [ops.pack_lift_step(pre.weight, blk.params()) for pre, blk in src])


class wavelet_forward_v2(nn.Module):
    def __init__(self, P, U, resnet_coeff, liftingLevel, convBlockList, cfg, nh=0, nl=0):
        super().__init__()
        self.P = P
        self.U = U
        self.resnet_weight = resnet_coeff
        self.lifting_level = liftingLevel
        self.csize = cfg.clrch
        self.convBlock = convBlockList
        self.nh = nh
        self.nl = nl
        self.scale = cfg.scale
        # optional key: "tc" (conv2/conv3 on tcgen05, 3xTF32 split) | "fp32" (all layers on the FP32 FMA pipe)
        self.lift_precision = cfg.get("lift_precision", ops.DEFAULT_LIFT_PRECISION) if hasattr(cfg, "get") else getattr(cfg, "lift_precision", ops.DEFAULT_LIFT_PRECISION)
        self._cache = PackCache()

    def _blobs(self):
        return pack_steps(self._cache, self.P, self.U, self.convBlock)

    def _linear(self):
        return self.P[0].linearityFlag != 1

    def one_level_lifting(self, x):
        """x (B,1,h,w) -> LL, LH, HL, HH (B,1,h/2,w/2) (wavelet_forward_v2.py:26-54)."""
        ll, yh = self.level(x)
        return ll, yh[:, 0:1], yh[:, 1:2], yh[:, 2:3]

    def level(self, x, ll_out=None, yh_out=None):
        """Same, returning (LL, Yh=(B,3,h/2,w/2) [LH,HL,HH]) without slicing."""
        scale = 1 if self.scale == 1 else 0
        params = step_params(self.P, self.U, self.convBlock, self.nh, self.nl, x)
        if _autograd.needs_grad([x] + params):
            # training: forward on the fused kernels, backward by recomputation in torch (see _autograd.py)
            fast = lambda x, *ps: ops.lift_level_fwd(x, self._blobs(), self.resnet_weight, self._linear(), scale,
                                                     self.nh if scale else None, self.nl if scale else None, precision=self.lift_precision)
            ref = lambda x, *ps: _torch_ref.lift_level_fwd(x, ps, self.resnet_weight, self._linear(), scale)
            return _autograd.run(fast, ref, [x] + params)
        return ops.lift_level_fwd(x, self._blobs(), self.resnet_weight, self._linear(), scale,
                                  self.nh if scale else None, self.nl if scale else None, ll_out, yh_out, precision=self.lift_precision)

    def lifting_forward_row_2_stage_lifting(self, L, H):
        """Four lifting steps along dim 2 of the (B,1,n,m) halves (wavelet_forward_v2.py:58-81)."""
        blobs = self._blobs()
        B, C, n, m = L.shape
        Lv, Hv = L.reshape(B * C, n, m), H.reshape(B * C, n, m)
        Lo, Ho = torch.empty_like(Lv, memory_format=torch.contiguous_format), torch.empty_like(Hv, memory_format=torch.contiguous_format)
        lin = self._linear()
        ops.lift_step([(Lv, Hv, Ho)], blobs[0], 1.0, self.resnet_weight, lin, self.lift_precision)
        ops.lift_step([(Ho, Lv, Lo)], blobs[1], 1.0, self.resnet_weight, lin, self.lift_precision)
        ops.lift_step([(Lo, Ho, Ho)], blobs[2], 1.0, self.resnet_weight, lin, self.lift_precision)
        ops.lift_step([(Ho, Lo, Lo)], blobs[3], 1.0, self.resnet_weight, lin, self.lift_precision)
        Lo, Ho = Lo.view(B, C, n, m), Ho.view(B, C, n, m)
        if self.scale == 1:
            Ho = Ho * (lifting_coeff[4] + self.nh * 0.1)
            Lo = Lo * (lifting_coeff[5] + self.nl * 0.1)
        return Lo, Ho
