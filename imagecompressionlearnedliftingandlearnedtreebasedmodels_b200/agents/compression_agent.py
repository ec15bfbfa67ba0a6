"""``CompressionAgent`` (agents/compression_agent.py:12-53): skeleton kept importable.

The reference class drives a ``model(x) -> (x_hat, rate)`` network that owns an ``entropy`` sub-module with
``find_cdf_range / quantize_cdf / display``.  No such model ships in the reference's hot path -- its constructor leaves
``self.model = None`` (:15) -- so only the contract of the two loops is mirrored: whoever plugs a model in gets the
reference's call order (range search before every training step, CDF quantisation once before validation, the
validation score ``1 / loss`` of the last batch, ``inf`` when no CDF range exists).
"""
import math

import torch

from ..graphs.losses.rate_dist import TrainRDLoss, ValidRDLoss


class CompressionAgent:
    def __init__(self, config, data_loader=None, device=None):
        self.config = config
        self.device = torch.device("cuda:0" if device is None else device)
        self.model = self.postprocess = self.optimizer = None
        self.data_loader = data_loader
        self.train_loss = TrainRDLoss(config.lambda_)
        self.valid_loss = ValidRDLoss(config.lambda_)
        try:
            self.lr = config.learning_rate
        except (AttributeError, KeyError):
            self.lr = 1e-4
        self.current_iteration = 0

    def _require(self, **parts):
        missing = [name for name, part in parts.items() if part is None]
        if missing:
            raise RuntimeError("CompressionAgent: set " + ", ".join(missing) + " first (the reference leaves them None)")

    def _batches(self):
        for x in self.data_loader:
            yield x.to(self.device)

    def train_one_epoch(self):
        self._require(model=self.model, optimizer=self.optimizer, data_loader=self.data_loader)
        self.model.train()
        for x in self._batches():
            self.model.entropy.find_cdf_range()
            self.optimizer.zero_grad()
            loss = self.train_loss(x, *self.model(x))[0]
            loss.backward()
            self.optimizer.step()
            self.current_iteration += 1

    @torch.no_grad()
    def validate(self):
        self._require(model=self.model, data_loader=self.data_loader)
        self.model.eval()
        entropy = self.model.entropy
        if not entropy.find_cdf_range():
            return math.inf
        entropy.quantize_cdf()
        score = math.inf
        for x in self._batches():
            loss = self.valid_loss(x, *self.model(x))[0]
            score = 1 / loss.item()
        return score
