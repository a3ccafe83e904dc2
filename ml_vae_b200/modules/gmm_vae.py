"""Drop-in for the reference's modules/gmm_vae.py:8-67 (GMMVAE), SURVEY.md section 8f-3.

Same constructor kwargs (fc_sizes, latent_size, num_components), checkpoint keys (fc.0.blocks.*, prior_mean_fc.*,
prior_log_var_fc.*, mean_fc.*, log_var_fc.*, gmm_weight_fc.*) and forward() dict.  The five heads read the trunk output
through one stacked GEMM; reparameterisation + KL against the learned prior is one fused kernel forward and backward
(csrc/latent_loss.cu, gmm_reparam_kl_*).  ``eps=`` / ``gumbels=`` inject the two noise sources for parity tests; left
None, eps is the Philox stream and the Gumbel noise comes from torch like in the reference (gmm_vae.py:31).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from ..dense import linear, linear_chain
from ._params import attach, torch_default_linear
from .vanilla_vae import default_stream_seed


def gumbel_softmax_hard(logits, tau: float, gumbels=None):
    """F.gumbel_softmax(logits, tau, hard=True) with optional injected Gumbel(0,1) noise (straight-through)."""
    if gumbels is None:
        return F.gumbel_softmax(logits, tau=tau, hard=True)
    y_soft = ((logits + gumbels.to(logits.dtype)) / tau).softmax(-1)
    index = y_soft.max(-1, keepdim=True)[1]
    y_hard = torch.zeros_like(logits).scatter_(-1, index, 1.0)
    return y_hard - y_soft.detach() + y_soft


class GMMVAE(nn.Module):
    HEADS = ("prior_mean_fc", "prior_log_var_fc", "mean_fc", "log_var_fc")

    def __init__(self, fc_sizes, latent_size, num_components, seed: int = None):
        super().__init__()
        self.fc_sizes = [int(s) for s in fc_sizes]
        self.latent_size, self.num_components = int(latent_size), int(num_components)
        for i in range(len(self.fc_sizes) - 1):
            w, b = torch_default_linear(self.fc_sizes[i], self.fc_sizes[i + 1])
            attach(self, f"fc.0.blocks.{2 * i}.weight", w)
            attach(self, f"fc.0.blocks.{2 * i}.bias", b)
        for head in self.HEADS:                                   # the reference's construction order (gmm_vae.py:17-22)
            w, b = torch_default_linear(self.fc_sizes[-1], self.latent_size * self.num_components)
            attach(self, f"{head}.weight", w)
            attach(self, f"{head}.bias", b)
        w, b = torch_default_linear(self.fc_sizes[-1], self.num_components)
        attach(self, "gmm_weight_fc.weight", w)
        attach(self, "gmm_weight_fc.bias", b)
        self.seed, self.calls = (default_stream_seed() if seed is None else int(seed)), 0

    def _trunk(self):
        blocks = self.fc._modules["0"].blocks._modules
        n = len(self.fc_sizes) - 1
        return [blocks[str(2 * i)].weight for i in range(n)], [blocks[str(2 * i)].bias for i in range(n)]

    def forward(self, feats, eps=None, gumbels=None):
        h = linear_chain(feats, *self._trunk(), end_activation=True)
        nl = self.latent_size * self.num_components
        w = torch.cat([getattr(self, n).weight for n in self.HEADS] + [self.gmm_weight_fc.weight], 0)
        b = torch.cat([getattr(self, n).bias for n in self.HEADS] + [self.gmm_weight_fc.bias], 0)
        o = linear(h, w, b)                                      # one pass over h for all five heads
        prior_mean, prior_log_var, mean, log_var = (o[..., i * nl:(i + 1) * nl].contiguous() for i in range(4))
        gmm_weight = gumbel_softmax_hard(o[..., 4 * nl:].float(), 0.1, gumbels).to(o.dtype)
        offset = self.calls
        self.calls += 1
        z, kl = ops.gmm_reparam_kl(mean, log_var, prior_mean, prior_log_var, eps=eps, seed=self.seed, offset=offset)
        return {"prior_mean": prior_mean, "prior_log_var": prior_log_var, "mean": mean, "log_var": log_var,
                "sampled_h": z, "gmm_weight": gmm_weight, "loss": kl}
