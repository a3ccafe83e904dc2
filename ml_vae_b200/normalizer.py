"""Drop-in for ``speechbrain.processing.features.InputNormalization(norm_type='global')`` as
declared at models/test_vanilla_vae/model.yaml:14-15 and called at
models/test_vanilla_vae/model.py:24-25 [arithmetic: SB-recall, SpeechBrain 0.5.x]:

  per utterance: mean and unbiased std over the round(len * T) valid frames (std floored at
  1e-10), averaged over the batch; running global statistics updated with weight
  1 / (count + 1) while epoch < update_until_epoch (3); output (x - glob_mean) / glob_std.

SpeechBrain loops over the batch in python with an .int() sync per utterance; this version is
sync-free (statistics stay on the device) and batched.  SURVEY.md section 8f-2 ("next" row).
"""
from __future__ import annotations

import torch

from . import _lib as L


class InputNormalization(torch.nn.Module):
    def __init__(self, mean_norm=True, std_norm=True, norm_type="global", avg_factor=None,
                 requires_grad=False, update_until_epoch=3):
        super().__init__()
        if norm_type != "global" or not (mean_norm and std_norm) or avg_factor is not None or requires_grad:
            raise NotImplementedError("only InputNormalization(norm_type='global') with SpeechBrain defaults "
                                      "(the reference's configuration, model.yaml:14-15) is implemented")
        self.update_until_epoch = update_until_epoch
        self.eps = 1e-10
        self.count = 0
        self.register_buffer("glob_mean", torch.zeros(0), persistent=False)
        self.register_buffer("glob_std", torch.zeros(0), persistent=False)

    @torch.no_grad()
    def batch_stats(self, x: torch.Tensor, lens: torch.Tensor):
        B, T, D = x.shape
        n = torch.round(lens.to(x.device).float() * T).clamp_(min=0, max=T)            # (B,)
        m = (torch.arange(T, device=x.device)[None, :] < n[:, None]).to(torch.float32)[..., None]
        xf = x.float()
        mean = (xf * m).sum(1) / n[:, None]
        var = (((xf - mean[:, None, :]) * m) ** 2).sum(1) / (n[:, None] - 1)
        std = var.sqrt().clamp_(min=self.eps)
        return mean.mean(0), std.mean(0)

    @torch.no_grad()
    def forward(self, x, lengths, epoch=0):
        L.require_cuda(x)
        if self.training:
            cm, cs = self.batch_stats(x, lengths)
            if self.count == 0:
                self.glob_mean, self.glob_std = cm, cs
            elif epoch < self.update_until_epoch:
                w = 1.0 / (self.count + 1)
                self.glob_mean = (1 - w) * self.glob_mean + w * cm
                self.glob_std = (1 - w) * self.glob_std + w * cs
            self.count += 1
        elif self.count == 0:
            raise RuntimeError("InputNormalization(global) used in eval mode before any statistics were seen")
        return ((x.float() - self.glob_mean) / self.glob_std).to(x.dtype)
