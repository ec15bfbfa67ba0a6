import torch
import torch.nn as nn

from oracle.thirdparty import lower_bound


class LowerBound(nn.Module):
    """compressai.ops.LowerBound: buffer ``bound`` + max with the custom gradient."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return lower_bound(x, self.bound)
