"""Rate-distortion losses (reference: graphs/losses/rate_dist.py:14-93).

bpp = sum(self-information) / (B*H*W): the reference writes it as ``sum / numel(x) * 3``.
The sums over the self-information tensors can also be taken inside the rate kernels
(``LiftingBasedDWTNetWrapper.set_bit_accumulator``); these classes keep the reference's
tensor-in / scalar-out surface.
"""
import torch
from torch import nn
import torch.nn.functional as F


class TrainRDLoss(nn.Module):
    def __init__(self, lambda_):
        super().__init__()
        self.mse_loss = nn.MSELoss(reduction='mean')
        self.lambda_ = lambda_

    def forward(self, x, x_hat, rate):
        self.mse = self.mse_loss(x, x_hat)
        self.rate = torch.sum(rate) / torch.numel(x) * 3
        self.loss = self.rate + self.lambda_ * self.mse
        return self.loss, self.mse, self.rate

    def forward2(self, x, x_hat, rate1, rate2):
        self.mse = self.mse_loss(x, x_hat)
        self.rate1 = torch.sum(rate1) / torch.numel(x) * 3
        self.rate2 = torch.sum(rate2) / torch.numel(x) * 3
        self.loss = self.rate1 + self.rate2 + self.lambda_ * self.mse
        return self.loss, self.mse, self.rate1, self.rate2

    def forward3(self, x, x_hat, rate1, rate2list):
        self.mse = self.mse_loss(x, x_hat)
        self.rate1 = torch.sum(rate1) / torch.numel(x) * 3
        self.rate2 = 0
        for i in range(len(rate2list)):
            self.rate2 += torch.sum(rate2list[i]) / torch.numel(x) * 3
        self.loss = self.rate1 + self.rate2 + self.lambda_ * self.mse
        return self.loss, self.mse, self.rate1, self.rate2


class TrainDLoss(TrainRDLoss):
    """Distortion-only warm-up loss (:45-72): rates are still reported."""

    def forward(self, x, x_hat, rate):
        self.mse = self.mse_loss(x, x_hat)
        self.rate = torch.sum(rate) / torch.numel(x) * 3
        self.loss = 0 + self.lambda_ * self.mse
        return self.loss, self.mse, self.rate

    def forward2(self, x, x_hat, rate1, rate2):
        self.mse = self.mse_loss(x, x_hat)
        self.rate1 = torch.sum(rate1) / torch.numel(x) * 3
        self.rate2 = torch.sum(rate2) / torch.numel(x) * 3
        self.loss = 0 + 0 + self.lambda_ * self.mse
        return self.loss, self.mse, self.rate1, self.rate2

    def forward3(self, x, x_hat, rate1, rate2list):
        self.mse = self.mse_loss(x, x_hat)
        self.rate1 = torch.sum(rate1) / torch.numel(x) * 3
        self.rate2 = 0
        for i in range(len(rate2list)):
            self.rate2 += torch.sum(rate2list[i]) / torch.numel(x) * 3
        self.loss = 0 + 0 + self.lambda_ * self.mse
        return self.loss, self.mse, self.rate1, self.rate2


class ValidRDLoss(nn.Module):
    def __init__(self, lambda_):
        super().__init__()
        self.lambda_ = lambda_

    def forward(self, x, x_hat, rate):
        self.mse = self.psnr(x, x_hat)
        if type(rate) == int:
            rate = torch.tensor([float(rate)])
        self.rate = torch.sum(rate, dtype=torch.float) / torch.numel(x) * 3
        self.loss = self.mse + self.rate * self.lambda_
        return self.loss, self.mse, self.rate

    def psnr(self, x, x_hat):
        mse = F.mse_loss(x_hat, x, reduction='none')
        mse = torch.mean(mse.view(mse.shape[0], -1), 1)
        psnr = -10 * torch.log10(mse)
        return torch.mean(psnr)
