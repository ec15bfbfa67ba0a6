"""Mean meters of the agent (reference: loggers/rate.py:50-151).  Host-side bookkeeping only; the
``state_dict`` layout ({loss, mse, rate, rate2, it, ep}) is what the reference's checkpoints carry under
``train_logger / trnit_logger / valid_logger / test_logger`` (agents/base.py:99-110), so those load unchanged."""
import logging
import math
from datetime import datetime


class RDLogger:
    _LABEL = {"tr": "  Train Epoch", "te": "   Test Epoch", "va": "  Valid Epoch", "it": "Train Itera"}

    def __init__(self):
        self.loss, self.mse, self.rate, self.rate2 = [], [], [], []
        self.current_iteration = 0
        self.current_epoch = 0
        self.logger = logging.getLogger("Loss")

    def __call__(self, loss, mse, rate, rate2=0):
        self.current_iteration += 1
        self.loss.append(loss)
        self.mse.append(mse)
        self.rate.append(rate)
        if rate2 > 0:
            self.rate2.append(rate2)

    append = __call__

    def reset(self):
        self.loss, self.mse, self.rate, self.rate2 = [], [], [], []

    def mean(self):
        """Means since the last call (and a reset), counted as one epoch."""
        self.current_epoch += 1
        avg = lambda v: sum(v) / len(v) if v else 0.0
        if not self.loss:
            raise RuntimeError("RDLogger.mean(): nothing was logged")
        out = avg(self.loss), avg(self.mse), avg(self.rate), avg(self.rate2)
        self.reset()
        return out

    def display(self, lr=0.0, typ="tr"):
        loss, mse, rate, rate2 = self.mean()
        psnr = 10.0 * math.log10(1.0 / mse) if mse > 0 else float("inf")
        rate_s = f"{rate:.3f}" if rate2 < 1e-6 else f"{rate:.3f}+{rate2:.3f}"
        lr_s = f"  (lr: {lr:.6f})" if typ in ("tr", "it") else ""
        self.logger.info(f"{self._LABEL.get(typ, typ)}: {self.current_epoch:3d}  RDLoss: {loss:.6f} MSE/PSNR: {mse:.6f}/{psnr:.2f} "
                         f"Rate: {rate_s}{lr_s} ({datetime.now().strftime('%H:%M:%S')})")
        return loss, mse, rate, rate2

    def state_dict(self):
        return {"loss": self.loss, "mse": self.mse, "rate": self.rate, "rate2": self.rate2,
                "it": self.current_iteration, "ep": self.current_epoch}

    def load_state_dict(self, info):
        self.loss, self.mse, self.rate, self.rate2 = info["loss"], info["mse"], info["rate"], info["rate2"]
        self.current_iteration, self.current_epoch = info["it"], info["ep"]
