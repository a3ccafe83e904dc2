"""B200 recipe: drop-in for the reference's src/models/test_vanilla_vae/model.py, discovered the
same way (prepare_experiment.py:47-49 imports ``models.<model_class>.model.SBModel``).

Copy (or symlink) this directory to ``src/models/b200_vanilla_vae/`` of a reference checkout and
run the reference's own entry point:

    python train.py config/run.yaml --dataset <ds> --model_class b200_vanilla_vae \
        --model_name b200 --model '!include:../models/b200_vanilla_vae/model.yaml'

compute_forward / compute_objectives keep the reference's structure (model.py:19-55) but ask the
drop-in modules for the length-masked means directly (one fused kernel each) instead of
materialising the unreduced (B, T, C) losses and reducing them with apply_lens_to_loss.
The loss keys ('kld_loss', 'recon_loss') are unchanged: they drive the weight lookup and logging
(md_model.py:189-213).
"""
from __future__ import annotations

try:                                    # inside a reference checkout with SpeechBrain installed
    from models.md_model import MDModel as _Base
    from utils.metric_stats.loss_metric_stats import LossMetricStats
except Exception:                       # standalone (tests, bench): same loop, no SpeechBrain
    from ml_vae_b200.brain import MiniBrain as _Base

    class LossMetricStats:              # utils/metric_stats/loss_metric_stats.py, device-resident
        def __init__(self, name):
            self.name, self.loss_list = name, []

        def append(self, loss):
            self.loss_list.append(loss.detach())          # no .cpu(): no per-loss host sync

        def summarize(self, field=None):
            import torch
            return {"loss": torch.stack(self.loss_list).mean().item()}


class SBModel(_Base):
    def on_stage_start(self, stage, epoch=None):
        if hasattr(super(), "on_stage_start"):
            super().on_stage_start(stage, epoch)
        self.stats_loggers["kld_loss_stats"] = LossMetricStats("kld_loss")
        self.stats_loggers["recon_loss_stats"] = LossMetricStats("recon_loss")

    def compute_forward(self, batch, stage):
        batch = batch.to(self.device)
        feats, feat_lens = batch["feat"]
        epoch = self.hparams.epoch_counter.current
        feats = self.hparams.normalizer(feats, feat_lens, epoch=epoch)
        enc = self.modules["encoder"]
        dec = self.modules["decoder"]
        enc.materialize_loss = dec.materialize_loss = False
        encoder_out = enc(feats, lens=feat_lens)
        decoder_out = dec(encoder_out["sampled_h"], feats, lens=feat_lens)
        return {"encoder_out": encoder_out, "decoder_out": decoder_out}

    def compute_objectives(self, predictions, batch, stage):
        losses = {"kld_loss": predictions["encoder_out"]["kld_loss"],
                  "recon_loss": predictions["decoder_out"]["recon_loss"]}
        return self.compute_and_save_losses(losses)
