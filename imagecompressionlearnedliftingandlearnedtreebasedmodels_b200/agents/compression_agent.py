"""``CompressionAgent`` (agents/compression_agent.py:12-53): skeleton kept importable.

The reference class drives a ``model(x) -> (x_hat, rate)`` network with an ``entropy`` sub-module exposing
``find_cdf_range / quantize_cdf / display``; no such model ships in the reference's hot path (``self.model = None``,
:15), so only the loop structure is mirrored."""
import math

import torch

from ..graphs.losses.rate_dist import TrainRDLoss


class CompressionAgent:
    def __init__(self, config, data_loader=None, device=None):
        self.config = config
        self.device = torch.device(device if device is not None else "cuda:0")
        self.model = None
        self.postprocess = None
        self.data_loader = data_loader
        self.train_loss = TrainRDLoss(config.lambda_)
        self.lr = config.learning_rate if "learning_rate" in dir(config) or (isinstance(config, dict) and "learning_rate" in config) else 1e-4
        self.optimizer = None
        self.current_iteration = 0

    def train_one_epoch(self):
        if self.model is None or self.optimizer is None or self.data_loader is None:
            raise RuntimeError("CompressionAgent: set model, optimizer and data_loader first (the reference leaves them None)")
        self.model.train()
        for x in self.data_loader:
            self.model.entropy.find_cdf_range()
            x = x.to(self.device)
            self.optimizer.zero_grad()
            x_hat, rate = self.model(x)
            loss, mse, rate = self.train_loss(x, x_hat, rate)
            loss.backward()
            self.optimizer.step()
            self.current_iteration += 1

    @torch.no_grad()
    def validate(self):
        if self.model is None or self.data_loader is None:
            raise RuntimeError("CompressionAgent: set model and data_loader first (the reference leaves them None)")
        self.model.eval()
        if not self.model.entropy.find_cdf_range():
            return math.inf
        self.model.entropy.quantize_cdf()
        loss = None
        for x in self.data_loader:
            x = x.to(self.device)
            x_hat, rate = self.model(x)
            loss, _, _ = self.train_loss(x, x_hat, rate)
        return 1 / loss.item()
