"""``P_block_v2`` -- the learned predict/update CNN of one lifting step.

Mirror of the reference class (graphs/layers/P_block_v2.py:7-55): same constructor,
same ``conv1..conv4`` parameters (so checkpoints load), forward through the fused
sm_100a lifting-step kernel.  Inside ``wavelet_forward_v2`` / ``wavelet_inverse_v2``
the block never runs on its own: the step kernel fuses it with the pre-filter and
the lifting update.
"""
import torch
import torch.nn as nn

from ... import ops

lifting_coeff = [-1.586134342059924, -0.052980118572961, 0.882911075530934, 0.443506852043971, 0.869864451624781,
                 1.149604398860241]  # bior4.4


class P_block_v2(nn.Module):
    def __init__(self, linearity_flag=1, csize=1, conv_filter_size=3, depth_scale=16):
        super().__init__()
        self.conv_filter_size = conv_filter_size
        self.padding = self.conv_filter_size // 2
        self.csize = csize
        k, p, d = self.conv_filter_size, self.padding, depth_scale
        self.conv1 = nn.Conv2d(1 * csize, d * csize, k, stride=1, padding=p)
        self.conv2 = nn.Conv2d(d * csize, d * csize, k, stride=1, padding=p)
        self.conv3 = nn.Conv2d(d * csize, d * csize, k, stride=1, padding=p)
        self.conv4 = nn.Conv2d(d * csize, 1 * csize, k, stride=1, padding=p)
        self.linearityFlag = linearity_flag
        self.nonLinearityFunction = nn.Tanh()
        self.lift_precision = ops.DEFAULT_LIFT_PRECISION

    def params(self):
        return {k: (getattr(self, k).weight, getattr(self, k).bias) for k in ("conv1", "conv2", "conv3", "conv4")}

    def forward(self, tmp):
        """conv1 -> tanh -> conv2 -> tanh -> conv3 + conv1 output -> conv4 (P_block_v2.py:40-55)."""
        identity = torch.tensor([0.0, 1.0, 0.0], device=tmp.device)
        blob = ops.pack_lift_step(identity, self.params())
        x = tmp.contiguous()
        B, C, h, w = x.shape
        out = torch.empty_like(x)
        v = x.view(B * C, h, w)
        ops.lift_step([(v, v, out.view(B * C, h, w))], blob, 0.0, 1.0, self.linearityFlag != 1, self.lift_precision)
        return out
