"""Drop-in for the reference's modules/vanilla_vae.py:9-45 (VanillaVAE).

Same constructor kwargs (fc_sizes, latent_size), same checkpoint keys
(fc.0.blocks.{0,2}.*, mean_fc.*, log_var_fc.*) and the same forward() dict
{'mean', 'log_var', 'sampled_h', 'loss'} with 'loss' the UNREDUCED (B, T, L) KL.

Differences that are the point of this package:
  * reparameterise + KL are one fused kernel (csrc/latent_loss.cu), forward and backward;
  * eps is the counter-based Philox stream (seed, per-call offset) regenerated in the
    backward instead of a stored torch.randn_like tensor; pass ``eps=`` to inject one
    (parity tests), or call ``set_seed``;
  * ``forward(feats, lens=...)`` additionally returns 'kld_loss', the length-masked mean
    (apply_lens_to_loss) computed inside the same kernel without materialising 'loss'
    when ``materialize_loss=False``.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import mlp_chain, ops
from ..dense import LEAKY_SLOPE, direct_chain, linear_chain, linear_direct
from ._params import attach, torch_default_linear

_DEFAULT_SEED = 123456      # run.yaml:2
_instances = 0              # VAE-family modules built so far in this process (construction order is deterministic)


def default_stream_seed() -> int:
    """Seed of the next module that was not given one: distinct per instance, so two latent blocks of one model
    (h_vae.py:17-18 builds a VanillaVAE and a GMMVAE) never draw the same eps -- the reference draws independent
    randn_like tensors (vanilla_vae.py:39, gmm_vae.py:53)."""
    global _instances
    _instances += 1
    return _DEFAULT_SEED + (_instances - 1)


class VanillaVAE(nn.Module):
    def __init__(self, fc_sizes, latent_size, seed: int = None, materialize_loss: bool = True):
        super().__init__()
        self.fc_sizes = [int(s) for s in fc_sizes]
        self.latent_size = int(latent_size)
        self.materialize_loss = materialize_loss
        for i in range(len(self.fc_sizes) - 1):
            w, b = torch_default_linear(self.fc_sizes[i], self.fc_sizes[i + 1])
            attach(self, f"fc.0.blocks.{2 * i}.weight", w)
            attach(self, f"fc.0.blocks.{2 * i}.bias", b)
        for head in ("mean_fc", "log_var_fc"):
            w, b = torch_default_linear(self.fc_sizes[-1], self.latent_size)
            attach(self, f"{head}.weight", w)
            attach(self, f"{head}.bias", b)
        self.seed = default_stream_seed() if seed is None else int(seed)
        self.calls = 0          # Philox offset: one fresh eps stream per forward
        self.offset_dev = None  # optional device int64[1] step counter used INSTEAD of `calls` (CUDA-graph replays)

    def set_seed(self, seed: int, calls: int = 0):
        self.seed, self.calls = int(seed), int(calls)

    def _trunk(self):
        blocks = self.fc._modules["0"].blocks._modules
        n = len(self.fc_sizes) - 1
        return [blocks[str(2 * i)].weight for i in range(n)], [blocks[str(2 * i)].bias for i in range(n)]

    # -- flat-arena binding (train_step.FlatArena) -------------------------------------------------------------
    def adjacent_param_groups(self):
        return [[self.mean_fc.weight, self.log_var_fc.weight], [self.mean_fc.bias, self.log_var_fc.bias]]

    def bind_arena(self, arena):
        ws, bs = self._trunk()
        trunk = [arena.linear_views([w], [b]) for w, b in zip(ws, bs)]
        heads = arena.linear_views([self.mean_fc.weight, self.log_var_fc.weight], [self.mean_fc.bias, self.log_var_fc.bias])
        self._direct = (trunk, heads) if heads is not None and all(v is not None for v in trunk) else None

    def _project_stacked(self, feats):
        """Arena-bound bf16 path: [mean | log_var] as ONE (B, T, 2L) tensor (one GEMM against the adjacent head weights), or None."""
        direct = getattr(self, "_direct", None)
        if direct is None or feats.dtype != torch.bfloat16:
            return None
        trunk, heads = direct
        if len(trunk) == 2 and mlp_chain.supported(trunk[0][0].shape[1], trunk[0][0].shape[0], trunk[1][0].shape[0]):
            h, = mlp_chain.chain2(feats, [trunk[0]], [trunk[1]], act_b=True)      # fused trunk (csrc/mlp_chain.cu)
        else:
            h = direct_chain(feats, trunk, end_activation=True)
        return linear_direct(h, heads)

    def project(self, feats):
        """feats (B, T, C) -> mean, log_var (B, T, L).  vanilla_vae.py:22-24."""
        ml = self._project_stacked(feats)
        if ml is not None:
            return ml[..., : self.latent_size].contiguous(), ml[..., self.latent_size:].contiguous()
        ws, bs = self._trunk()
        h = linear_chain(feats, ws, bs, end_activation=True)
        # both heads read h once: one GEMM against the stacked (2L, H) weight
        w = torch.cat([self.mean_fc.weight, self.log_var_fc.weight], 0)
        b = torch.cat([self.mean_fc.bias, self.log_var_fc.bias], 0)
        ml = linear_chain(h, [w], [b])
        mean, log_var = ml[..., : self.latent_size], ml[..., self.latent_size:]
        return mean.contiguous(), log_var.contiguous()

    def forward(self, feats, lens=None, eps=None):
        # with a device step counter the kernel adds it to `offset`; the host call count moves to the high word so that
        # repeated forwards between two counter increments (eval batches, gradient accumulation) still get fresh eps
        offset = (self.calls << 32) if self.offset_dev is not None else self.calls
        self.calls += 1
        ml = self._project_stacked(feats)
        if ml is not None:
            # the kernel reads both halves of the stacked projection in place and its backward writes one stacked gradient:
            # 'mean' / 'log_var' are views (no slice copies, no zero-fill + scatter + add of slice gradients in backward)
            mean, log_var = ml[..., : self.latent_size], ml[..., self.latent_size:]
            z, kl_elem, kl_mean = ops.reparam_kl_stacked(ml, lens=lens, eps=eps, seed=self.seed, offset=offset,
                                                         want_elem=self.materialize_loss, want_mean=lens is not None,
                                                         offset_dev=self.offset_dev)
        else:
            mean, log_var = self.project(feats)
            z, kl_elem, kl_mean = ops.reparam_kl(mean, log_var, lens=lens, eps=eps, seed=self.seed, offset=offset,
                                                 want_elem=self.materialize_loss, want_mean=lens is not None,
                                                 offset_dev=self.offset_dev)
        out = {"mean": mean, "log_var": log_var, "sampled_h": z, "loss": kl_elem}
        if lens is not None:
            out["kld_loss"] = kl_mean
        return out

    # kept for API parity with the reference class
    def reparameterize(self, mean, log_var, eps=None):
        offset = self.calls
        self.calls += 1
        return ops.reparam_kl(mean, log_var, eps=eps, seed=self.seed, offset=offset,
                              want_elem=False, want_mean=False)[0]

    def compute_kld_loss(self, mean, log_var):
        return ops.reparam_kl(mean, log_var, eps=torch.zeros_like(mean), want_elem=True, want_mean=False)[1]
