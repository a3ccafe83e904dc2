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


    # Checkpointer hooks (what speechbrain's @mark_as_saver / @mark_as_loader give its EpochCounter)
    def state_dict(self):
        return {"current": self.current}

    def load_state_dict(self, sd):
        self.current = int(sd["current"])


class Checkpointer:
    """The slice of ``speechbrain.utils.checkpoints.Checkpointer`` the reference touches
    (models/test_vanilla_vae/model.yaml:6-12, prepare_experiment.py:56, md_model.py:50-52,162-164):
    ``recoverables`` by name, ``add_recoverable``, ``save_and_keep_only(meta, max_keys, min_keys)``,
    ``recover_if_possible``.  Every recoverable is anything with state_dict()/load_state_dict() (modules,
    optimisers, EpochCounter, the device-resident InputNormalization); one directory per checkpoint, one
    ``<name>.ckpt`` (torch.save of the state_dict) per recoverable, ``meta.json`` beside them."""

    def __init__(self, checkpoints_dir, recoverables=None):
        import os
        self.dir = os.fspath(checkpoints_dir)
        self.recoverables = dict(recoverables or {})

    def add_recoverable(self, name, obj):
        self.recoverables[name] = obj

    def _list(self):
        import json, os
        out = []
        if os.path.isdir(self.dir):
            for d in sorted(os.listdir(self.dir)):
                m = os.path.join(self.dir, d, "meta.json")
                if d.startswith("CKPT+") and os.path.exists(m):
                    out.append((os.path.join(self.dir, d), json.load(open(m))))
        return out

    def save_checkpoint(self, meta=None):
        import json, os, time
        existing = self._list()
        path = os.path.join(self.dir, f"CKPT+{time.strftime('%Y-%m-%d+%H-%M-%S')}+{len(existing):02d}")
        os.makedirs(path, exist_ok=True)
        for name, obj in self.recoverables.items():
            torch.save(obj.state_dict(), os.path.join(path, f"{name}.ckpt"))
        m = {k: (float(v) if hasattr(v, "__float__") else v) for k, v in (meta or {}).items()}
        m["unixtime"] = time.time()
        with open(os.path.join(path, "meta.json"), "w") as f:
            json.dump(m, f)
        return path

    def save_and_keep_only(self, meta=None, max_keys=(), min_keys=(), num_to_keep=1):
        import shutil
        self.save_checkpoint(meta)
        ck = self._list()
        keep = set(p for p, _ in sorted(ck, key=lambda c: c[1]["unixtime"])[-num_to_keep:])
        for key in max_keys:
            c = [x for x in ck if key in x[1]]
            keep |= set(p for p, _ in sorted(c, key=lambda x: x[1][key])[-num_to_keep:])
        for key in min_keys:
            c = [x for x in ck if key in x[1]]
            keep |= set(p for p, _ in sorted(c, key=lambda x: x[1][key])[:num_to_keep])
        for p, _ in ck:
            if p not in keep:
                shutil.rmtree(p, ignore_errors=True)

    def recover_if_possible(self, max_key=None, min_key=None, device=None):
        import os
        ck = self._list()
        if not ck:
            return None
        if max_key is not None:
            ck = sorted([c for c in ck if max_key in c[1]], key=lambda c: c[1][max_key])
        elif min_key is not None:
            ck = sorted([c for c in ck if min_key in c[1]], key=lambda c: -c[1][min_key])
        else:
            ck = sorted(ck, key=lambda c: c[1]["unixtime"])
        path, meta = ck[-1]
        for name, obj in self.recoverables.items():
            f = os.path.join(path, f"{name}.ckpt")
            if os.path.exists(f):
                obj.load_state_dict(torch.load(f, map_location=device, weights_only=False))
        return path, meta


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
        if self.checkpointer is not None:                      # md_model.py:50-52
            for key, optimizer in self.optimizers.items():
                self.checkpointer.add_recoverable(key, optimizer)

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
