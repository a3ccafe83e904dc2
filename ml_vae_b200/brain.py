"""Minimal stand-in for the slice of ``speechbrain.Brain`` the reference's MDModel uses on the
training hot path, so the B200 recipe runs (and is tested) where SpeechBrain is not installed.

Mirrors, for one optimiser or several:
  models/md_model.py:20-52    init_optimizers   (hparams.optimizer / hparams.optimizers)
  models/md_model.py:54-88    fit_batch         (non-AMP branch: fwd, objectives, backward,
                                                 check_gradients, step for every optimiser, zero_grad)
  models/md_model.py:189-213  compute_and_save_losses
check_gradients follows SpeechBrain 0.5.x [SB-recall]: a non-finite loss skips the update
(up to nonfinite_patience times, then raises) and the global gradient norm is clipped to
max_grad_norm = 5.0.

With SpeechBrain installed, ``models/b200_vanilla_vae/model.py`` subclasses the reference's own
MDModel instead and this class is not used.
"""
from __future__ import annotations

import warnings
from types import SimpleNamespace

import torch

from .train_step import KLD_N_SAMPLES


class EpochCounter:
    """speechbrain.utils.epoch_loop.EpochCounter subset: iterate 1..limit, remember ``current``."""

    def __init__(self, limit):
        self.current, self.limit = 0, int(limit)

    def __iter__(self):
        return self

    def __next__(self):
        if self.current < self.limit:
            self.current += 1
            return self.current
        raise StopIteration


class Stage:
    TRAIN, VALID, TEST = "TRAIN", "VALID", "TEST"


class PaddedBatchLite(dict):
    """What the recipes touch of speechbrain's PaddedBatch: batch['feat'] -> (data, rel_lens), .to()."""

    def to(self, device):
        return PaddedBatchLite({k: tuple(t.to(device) if torch.is_tensor(t) else t for t in v)
                                if isinstance(v, tuple) else v for k, v in self.items()})


class MiniBrain:
    def __init__(self, modules=None, hparams=None, run_opts=None, checkpointer=None, label_encoder=None):
        self.modules = torch.nn.ModuleDict(modules or {})
        self.hparams = SimpleNamespace(**(hparams or {}))
        run_opts = run_opts or {}
        self.device = torch.device(run_opts.get("device", "cuda:0"))
        self.max_grad_norm = run_opts.get("max_grad_norm", 5.0)
        self.nonfinite_patience = run_opts.get("nonfinite_patience", 3)
        self.nonfinite_count = 0
        self.checkpointer, self.label_encoder = checkpointer, label_encoder
        self.stats_loggers = {}
        self.modules.to(self.device)
        self.init_optimizers()

    # md_model.py:20-52
    def init_optimizers(self):
        if hasattr(self.hparams, "optimizers"):
            info = self.hparams.optimizers
            if isinstance(info, list):
                info = {f"optimizer_{i}": o for i, o in enumerate(info)}
        elif hasattr(self.hparams, "optimizer"):
            info = {"optimizer": self.hparams.optimizer}
        else:
            raise ValueError("No optimizers defined.")
        self.optimizers = {}
        for key, o in info.items():
            if isinstance(o, dict):
                params = ([p for name in o["modules"] for p in self.modules[name].parameters()]
                          if "modules" in o else self.modules.parameters())
                self.optimizers[key] = o["opt_class"](params)
            else:
                self.optimizers[key] = o(self.modules.parameters())

    # SpeechBrain Brain.check_gradients [SB-recall]
    def check_gradients(self, loss):
        if not torch.isfinite(loss):
            self.nonfinite_count += 1
            warnings.warn(f"Loss is {loss}.")
            if self.nonfinite_count > self.nonfinite_patience:
                raise ValueError("Loss is not finite and patience is exhausted.")
            return False
        torch.nn.utils.clip_grad_norm_((p for p in self.modules.parameters()), self.max_grad_norm)
        return True

    # md_model.py:78-88
    def fit_batch(self, batch):
        optimizers = list(self.optimizers.values())
        outputs = self.compute_forward(batch, Stage.TRAIN)
        loss = self.compute_objectives(outputs, batch, Stage.TRAIN)
        loss.backward()
        if self.check_gradients(loss):
            for o in optimizers:
                o.step()
        for o in optimizers:
            o.zero_grad()
        return loss.detach().cpu()

    # md_model.py:189-213
    def compute_and_save_losses(self, losses):
        loss = 0
        for key, value in losses.items():
            wkey = key.replace("_loss", "_weight")
            weight = getattr(self.hparams, wkey, "none")
            if weight == "none":
                warnings.warn(f"{wkey} not found, use 1 as default")
                weight = 1
            if "_kld" in wkey:
                weight /= (KLD_N_SAMPLES / self.hparams.batch_size)
            loss += weight * value
            logger = self.stats_loggers.get(key + "_stats")
            if logger is not None:
                logger.append(value)
        return loss
