"""``BaseAgent`` (reference: agents/base.py:13-187): run-mode dispatch, epoch loop, checkpoint save / resume.

Kept verbatim in *semantics* (SURVEY.md section 5): the checkpoint is one dict {epoch, iteration, best_valid_loss,
state_dict, optimizer, scheduler, train_logger, trnit_logger, valid_logger, test_logger} written to
``<checkpoint_dir>/checkpoint.pth.tar`` and copied to ``model_best.pth.tar`` when best; ``load_checkpoint`` restores the
model and the loggers but not the optimizer / scheduler (:74-75 are commented out upstream) and skips silently when the
file does not exist; ``run()`` swallows ``KeyboardInterrupt``, re-raises ``AssertionError`` untouched and, on any other
exception, saves a checkpoint before re-raising (:148-154).  What changes: the device follows ``gpu_device`` (or
``LOCAL_RANK`` under torchrun) instead of the reference's hard ``"cuda"`` / CPU-in-test-mode switch -- this package has
no CPU path -- and the SMTP mailer import is gone.
"""
import logging
import os
import shutil

import torch


class BaseAgent:
    def __init__(self, config, device=None):
        self.config = config
        self.logger = logging.getLogger("Agent")
        self.best_valid_loss = float("inf")
        self.current_epoch = 0
        self.current_iteration = 0
        if device is None:
            idx = int(os.environ.get("LOCAL_RANK", self._get("gpu_device", 0)))
            device = torch.device("cuda", idx)
        self.device = torch.device(device)
        self.manual_seed = self._get("seed", 1337)
        self.lr = self._get("learning_rate", 1e-4)
        if self.device.type == "cuda" and torch.cuda.is_available():
            torch.cuda.manual_seed(self.manual_seed)
            torch.cuda.set_device(self.device)

    def _get(self, key, default):
        """Optional config key (EasyDict raises AttributeError, dict-backed configs KeyError)."""
        try:
            return getattr(self.config, key)
        except (AttributeError, KeyError):
            return default

    # ---- to be provided by the concrete agent ----
    def train_one_epoch(self):
        raise NotImplementedError

    def validate(self):
        raise NotImplementedError

    def validate_recu_reco(self):
        raise NotImplementedError

    def test(self):
        raise NotImplementedError

    # ---- checkpoints (:64-128) ----
    def _loggers(self):
        return {k: getattr(self, k) for k in ("train_logger", "trnit_logger", "valid_logger", "test_logger")}

    def load_checkpoint(self, filename):
        path = self._get("checkpoint_dir", "") + filename
        try:
            self.logger.info("Loading checkpoint '{}'".format(path))
            checkpoint = torch.load(path, map_location=self.device, weights_only=False)
        except OSError:
            self.logger.info("No checkpoint exists from '{}'. Skipping...".format(self._get("checkpoint_dir", "")))
            self.logger.info("**First time to train**")
            return False
        self.current_epoch = checkpoint["epoch"]
        self.current_iteration = checkpoint["iteration"]
        self.best_valid_loss = checkpoint["best_valid_loss"]
        self.model.load_state_dict(checkpoint["state_dict"])
        for name, lg in self._loggers().items():
            if name in checkpoint:
                lg.load_state_dict(checkpoint[name])
        self.model.to(self.device)
        self.logger.info("Checkpoint loaded successfully from '{}' at (epoch {}) at (iteration {})\n".format(
            self._get("checkpoint_dir", ""), checkpoint["epoch"], checkpoint["iteration"]))
        return True

    def save_checkpoint(self, filename="checkpoint.pth.tar", is_best=0):
        state = {"epoch": self.current_epoch, "iteration": self.current_iteration,
                 "best_valid_loss": self.best_valid_loss, "state_dict": self.model.state_dict(),
                 "optimizer": self.optimizer.state_dict(), "scheduler": self.scheduler.state_dict()}
        state.update({k: lg.state_dict() for k, lg in self._loggers().items()})
        ckpt_dir = self._get("checkpoint_dir", "")
        if ckpt_dir:
            os.makedirs(ckpt_dir, exist_ok=True)
        torch.save(state, ckpt_dir + filename)
        if is_best:
            shutil.copyfile(ckpt_dir + filename, ckpt_dir + "model_best.pth.tar")

    # ---- run modes (:130-168) ----
    def run(self):
        mode = self.config.mode
        try:
            if mode == "test":
                self.test()
            elif mode == "validate":
                self.validate()
            elif mode == "validate_recu_reco":
                self.validate_recu_reco()
            elif mode == "train":
                self.train()
            elif mode == "debug":
                with torch.autograd.detect_anomaly():
                    self.train()
            else:
                raise NameError("'" + str(mode) + "'" + " is not a valid training mode.")
        except KeyboardInterrupt:
            self.logger.info("You have entered CTRL+C.. Wait to finalize")
        except AssertionError:
            raise
        except Exception:
            self.save_checkpoint()
            raise

    def train(self):
        for epoch in range(self.current_epoch, self.config.max_epoch):
            self.current_epoch = epoch
            self.train_one_epoch()
            if not (self.current_epoch + 1) % self._get("validate_every", 1):
                valid_loss = self.validate()
                is_best = valid_loss < self.best_valid_loss
                if is_best:
                    self.best_valid_loss = valid_loss
                self.save_checkpoint(is_best=is_best)
            self.current_epoch += 1

    def finalize(self):
        self.logger.info("Please wait while finalizing the operation.. Thank you")
        if self.config.mode == "train":
            self.save_checkpoint()
