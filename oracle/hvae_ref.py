"""CPU oracle: functional restatement of the reference's GMM-VAE / hierarchical-VAE latent family
(SURVEY.md section 8f-3).  TEST INFRASTRUCTURE ONLY.  PINNED: oracle/gen_golden.py checks it against the
reference's own modules.gmm_vae.GMMVAE / modules.h_vae.HierarchicalVAE and stores their outputs in
tests/golden/hvae_small.npz.

  gumbel_softmax_st   torch.nn.functional.gumbel_softmax(tau, hard=True) with the Gumbel noise supplied
                      (called at /root/reference/src/modules/gmm_vae.py:31 with tau=0.1)
  gmm_kld_elementwise /root/reference/src/modules/gmm_vae.py:58-67
  apply_weight        /root/reference/src/utils/data_utils.py:32-64   (bmm of (M,1,N) x (M,N,C))
  gmm_forward         /root/reference/src/modules/gmm_vae.py:24-50
  hvae_forward        /root/reference/src/modules/h_vae.py:22-72
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import vae_ref

GMM_EPS = 1e-5      # gmm_vae.py:62


def gumbel_softmax_st(logits, gumbels, tau: float = 0.1):
    """Straight-through hard Gumbel-softmax, torch's formulation, noise injected."""
    y_soft = ((logits + gumbels) / tau).softmax(-1)
    index = y_soft.max(-1, keepdim=True)[1]
    y_hard = torch.zeros_like(logits).scatter_(-1, index, 1.0)
    return y_hard - y_soft.detach() + y_soft


def gmm_kld_elementwise(prior_mean, prior_log_var, mean, log_var):
    return -0.5 * (1 + log_var - prior_log_var
                   - (log_var.exp() + (mean - prior_mean) ** 2) / (prior_log_var.exp() + GMM_EPS))


def apply_weight(x, weight):
    """x (B,T,N*C) or (B,T,N,C), weight (B,T,N) -> (B,T,C): sum_n weight[n] * x[n]."""
    B, T, N = weight.shape
    x = x.reshape(B, T, N, -1)
    return torch.bmm(weight.reshape(B * T, 1, N), x.reshape(B * T, N, -1)).reshape(B, T, -1)


def gmm_forward(params: dict, feats, eps, gumbels):
    h = F.leaky_relu(vae_ref.fc_stack(params, "fc.0.blocks", feats), vae_ref.LEAKY_SLOPE)
    lin = lambda n: F.linear(h, params[f"{n}.weight"], params[f"{n}.bias"])
    prior_mean, prior_log_var, mean, log_var = lin("prior_mean_fc"), lin("prior_log_var_fc"), lin("mean_fc"), lin("log_var_fc")
    w = gumbel_softmax_st(lin("gmm_weight_fc"), gumbels, 0.1)
    return {"prior_mean": prior_mean, "prior_log_var": prior_log_var, "mean": mean, "log_var": log_var,
            "sampled_h": vae_ref.reparameterize(mean, log_var, eps), "gmm_weight": w,
            "loss": gmm_kld_elementwise(prior_mean, prior_log_var, mean, log_var)}


def hvae_forward(params: dict, feats, pi, eps_vanilla, eps_gmm, gumbels):
    """params keyed like HierarchicalVAE.state_dict(): 'vanilla_vae.*', 'gmm_vae.*'."""
    sub = lambda pre: {k[len(pre):]: v for k, v in params.items() if k.startswith(pre)}
    van = vae_ref.encoder_forward(sub("vanilla_vae."), feats, eps_vanilla)
    gmm = gmm_forward(sub("gmm_vae."), feats, eps_gmm, gumbels)
    w = gmm["gmm_weight"]
    mix = {k: apply_weight(gmm[k], w) for k in ("mean", "log_var", "sampled_h", "loss")}
    out = {}
    for name, vk in (("mean", "mean"), ("log_var", "log_var"), ("sampled_h", "sampled_h"), ("kld", "loss")):
        stacked = torch.stack([van[vk], mix[vk]], dim=2)
        out[name] = apply_weight(stacked, pi)
    return {"gmm_weight": w, "mean": out["mean"], "log_var": out["log_var"], "sampled_h": out["sampled_h"],
            "losses": {"vae_kld_loss": out["kld"]}}
