"""Drop-in for the reference's modules/h_vae.py:12-72 (HierarchicalVAE), SURVEY.md section 8f-3.

Same constructor kwargs, checkpoint keys (vanilla_vae.*, gmm_vae.*) and forward(feats, pi) dict
{'gmm_weight','mean','log_var','sampled_h','losses':{'vae_kld_loss'}}.  The eight apply_weight calls of the reference
(M tiny bmm problems each) run on one CUDA kernel (ops.apply_weight, forward and backward incl. the gradient with
respect to the weights that the straight-through Gumbel-softmax needs).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from .gmm_vae import GMMVAE
from .vanilla_vae import VanillaVAE


class HierarchicalVAE(nn.Module):
    def __init__(self, fc_sizes, latent_size, num_components, seed: int = None):
        super().__init__()
        # two independent eps streams (h_vae.py:17-18 draws two randn_like tensors): seed / seed + 1 when given,
        # else the per-instance defaults
        self.vanilla_vae = VanillaVAE(fc_sizes, latent_size, seed=seed)                                   # correct pronunciation
        self.gmm_vae = GMMVAE(fc_sizes, latent_size, num_components, seed=None if seed is None else seed + 1)  # mispronunciation

    def forward(self, feats, pi, eps_vanilla=None, eps_gmm=None, gumbels=None):
        van = self.vanilla_vae(feats, eps=eps_vanilla)
        gmm = self.gmm_vae(feats, eps=eps_gmm, gumbels=gumbels)
        w = gmm["gmm_weight"]
        pi = pi.to(feats.dtype)
        out = {}
        for name, key in (("mean", "mean"), ("log_var", "log_var"), ("sampled_h", "sampled_h"), ("kld", "loss")):
            mixed = ops.apply_weight(gmm[key], w)                                 # (B, T, C)
            out[name] = ops.apply_weight(torch.stack([van[key], mixed], dim=2), pi)
        return {"gmm_weight": w, "mean": out["mean"], "log_var": out["log_var"], "sampled_h": out["sampled_h"],
                "losses": {"vae_kld_loss": out["kld"]}}
