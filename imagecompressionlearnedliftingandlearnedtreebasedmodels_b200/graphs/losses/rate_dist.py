"""Rate-distortion objectives with the reference's surface (graphs/losses/rate_dist.py:14-93).

One implementation serves the three call shapes: ``forward`` (one rate tensor), ``forward2`` (two), ``forward3``
(LL rate + a list of detail-level rates, the shape ``LiftingBasedDWTAgent`` uses).  The reference writes bits per pixel
as ``sum(self_information) / numel(x) * 3`` (:22,29-30,37-40), i.e. total bits / (B*H*W) for a 3-channel ``x``; that
expression is kept literally (``_bpp``) so the numbers agree to the last bit.  ``TrainDLoss`` (:45-72) is the
distortion-only warm-up objective: same reported rates, loss = lambda * mse.  The per-subband sums can also be taken
inside the rate kernels (``LiftingBasedDWTNetWrapper.set_bit_accumulator``); these classes are the tensor-in /
scalar-out form.  Results are also left on the module (``.loss``, ``.mse``, ``.rate`` / ``.rate1`` / ``.rate2``) like the
reference does.
"""
import torch
from torch import nn
import torch.nn.functional as F


def _bpp(x, bits):
    """Reference arithmetic for one tensor of self-informations: sum / numel(x) * 3."""
    return torch.sum(bits) / torch.numel(x) * 3


class TrainRDLoss(nn.Module):
    RATE_WEIGHT = 1          # 0 in TrainDLoss: the rate terms are reported but do not enter the loss

    def __init__(self, lambda_):
        super().__init__()
        self.lambda_ = lambda_
        self.mse_loss = nn.MSELoss(reduction='mean')

    def _objective(self, rates):
        # the reference adds its rate terms left to right before the distortion term; with weight 0 it writes "0 + ..."
        total = 0
        for r in rates:
            total = total + (r if self.RATE_WEIGHT else 0)
        self.loss = total + self.lambda_ * self.mse
        return self.loss

    def forward(self, x, x_hat, rate):
        self.mse = self.mse_loss(x, x_hat)
        self.rate = _bpp(x, rate)
        return self._objective([self.rate]), self.mse, self.rate

    def forward2(self, x, x_hat, rate1, rate2):
        self.mse = self.mse_loss(x, x_hat)
        self.rate1, self.rate2 = _bpp(x, rate1), _bpp(x, rate2)
        return self._objective([self.rate1, self.rate2]), self.mse, self.rate1, self.rate2

    def forward3(self, x, x_hat, rate1, rate2list):
        self.mse = self.mse_loss(x, x_hat)
        self.rate1 = _bpp(x, rate1)
        acc = 0
        for level_bits in rate2list:          # finest level first, accumulated in list order like the reference
            acc = acc + _bpp(x, level_bits)
        self.rate2 = acc
        return self._objective([self.rate1, self.rate2]), self.mse, self.rate1, self.rate2


class TrainDLoss(TrainRDLoss):
    """Distortion-only objective used until ``loss_switch_thr`` (agents/liftingDWT_agent.py:55-58)."""
    RATE_WEIGHT = 0


class ValidRDLoss(nn.Module):
    """Validation objective of ``CompressionAgent`` (:74-93): mean per-image PSNR + lambda * bpp."""

    def __init__(self, lambda_):
        super().__init__()
        self.lambda_ = lambda_

    @staticmethod
    def psnr(x, x_hat):
        per_image = F.mse_loss(x_hat, x, reduction='none').flatten(1).mean(dim=1)
        return (-10 * torch.log10(per_image)).mean()

    def forward(self, x, x_hat, rate):
        self.mse = self.psnr(x, x_hat)               # (sic: the reference stores the PSNR under this name)
        bits = torch.tensor([float(rate)]) if isinstance(rate, int) else rate
        self.rate = torch.sum(bits, dtype=torch.float) / torch.numel(x) * 3
        self.loss = self.mse + self.rate * self.lambda_
        return self.loss, self.mse, self.rate
