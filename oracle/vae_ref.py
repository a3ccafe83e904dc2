"""CPU oracle: functional restatement of the reference's VAE latent block.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED: oracle/gen_golden.py
checks every function here against the reference's own modules imported from
/root/reference/src and stores their outputs under tests/golden/.

Weights travel as plain ``dict[str, Tensor]`` keyed by the reference's
``state_dict`` names, so a reference checkpoint feeds the oracle directly.

  fc_stack            /root/reference/src/modules/fc_block.py:4-21
  encoder_forward     /root/reference/src/modules/vanilla_vae.py:21-35
  reparameterize      /root/reference/src/modules/vanilla_vae.py:37-40
  kld_elementwise     /root/reference/src/modules/vanilla_vae.py:42-45
  decoder_forward     /root/reference/src/modules/decoder.py:21-35
  recon_elementwise   /root/reference/src/modules/decoder.py:37-53
  length_mask         speechbrain.nnet.losses.length_to_mask [SB-recall] as used by
  masked_reduce       /root/reference/src/utils/data_utils.py:67-104
  weighted_total      /root/reference/src/models/md_model.py:189-213
  recipe_loss         /root/reference/src/models/test_vanilla_vae/model.py:19-55
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.01          # nn.LeakyReLU() default used throughout fc_block.py
RECON_EPS = 1e-5            # decoder.py:41
KLD_N_SAMPLES = 2249        # md_model.py:199


# ----------------------------------------------------------------------------
# dense stacks
# ----------------------------------------------------------------------------
def fc_stack(params: dict, prefix: str, x: torch.Tensor, end_activation: bool = False):
    """Linear -> LeakyReLU repeated, last Linear bare (fc_block.py:9-16).
    ``prefix`` is e.g. 'fc.0.blocks' and layers sit at even indices 0, 2, 4..."""
    idx = sorted(int(k[len(prefix) + 1:].split(".")[0]) for k in params
                 if k.startswith(prefix + ".") and k.endswith(".weight"))
    for j, i in enumerate(idx):
        x = F.linear(x, params[f"{prefix}.{i}.weight"], params[f"{prefix}.{i}.bias"])
        if j + 1 < len(idx) or end_activation:
            x = F.leaky_relu(x, LEAKY_SLOPE)
    return x


def reparameterize(mean, log_var, eps):
    """vanilla_vae.py:37-40 with eps supplied instead of drawn."""
    return eps * torch.exp(0.5 * log_var) + mean


def kld_elementwise(mean, log_var):
    """vanilla_vae.py:43, unreduced (B, T, L)."""
    return -0.5 * (1 + log_var - mean.pow(2) - log_var.exp())


def encoder_forward(params: dict, feats, eps):
    """VanillaVAE.forward (vanilla_vae.py:21-35) with injected eps."""
    h = F.leaky_relu(fc_stack(params, "fc.0.blocks", feats), LEAKY_SLOPE)
    mean = F.linear(h, params["mean_fc.weight"], params["mean_fc.bias"])
    log_var = F.linear(h, params["log_var_fc.weight"], params["log_var_fc.bias"])
    return {"mean": mean, "log_var": log_var,
            "sampled_h": reparameterize(mean, log_var, eps),
            "loss": kld_elementwise(mean, log_var)}


def recon_elementwise(mean, log_var, target, loss_type: str = "likelihood"):
    """decoder.py:37-53.  log(2*pi) is evaluated in float32 like the reference
    (torch.log(2 * torch.tensor(np.pi))) and then promoted."""
    if loss_type == "likelihood":
        log_2pi = torch.log(2 * torch.tensor(np.pi))
        return 0.5 * (log_2pi + log_var + (target - mean) ** 2 / (torch.exp(log_var) + RECON_EPS))
    if loss_type == "mse":
        return (target - mean) ** 2
    raise ValueError(f"Invalid loss type: {loss_type}")


def bilstm(params: dict, x, hidden: int, num_layers: int, drop_masks=None, drop_p: float = 0.0, torch_dropout: float = 0.0):
    """decoder.py:14-15,22: batch_first bidirectional LSTM.  ``drop_masks`` (list of num_layers - 1 boolean keep masks of
    shape (B, T, 2*hidden)) injects the inter-layer dropout nn.LSTM(dropout=p) applies in training mode to the output of
    every layer but the last: out * keep / (1 - p).  None = dropout disabled (eval mode / p = 0).  ``torch_dropout`` > 0
    runs the stacked LSTM in training mode with torch's own (irreproducible) dropout: timing runs only."""
    def layer_params(layer):
        flat = []
        for sfx in ("", "_reverse"):
            flat += [params[f"rnn.weight_ih_l{layer}{sfx}"], params[f"rnn.weight_hh_l{layer}{sfx}"],
                     params[f"rnn.bias_ih_l{layer}{sfx}"], params[f"rnn.bias_hh_l{layer}{sfx}"]]
        return flat
    z = x.new_zeros(2, x.shape[0], hidden)
    if drop_masks is None:
        flat = [p for layer in range(num_layers) for p in layer_params(layer)]
        zz = x.new_zeros(2 * num_layers, x.shape[0], hidden)
        out, _, _ = torch._VF.lstm(x, (zz, zz), flat, True, num_layers, float(torch_dropout), torch_dropout > 0, True, True)
        return out
    out = x
    for layer in range(num_layers):
        out, _, _ = torch._VF.lstm(out, (z, z), layer_params(layer), True, 1, 0.0, False, True, True)
        if layer + 1 < num_layers:
            scale = torch.tensor(1.0, dtype=torch.float32) / (1.0 - torch.tensor(drop_p, dtype=torch.float32))
            out = out * drop_masks[layer].to(out.dtype) * scale.to(out.dtype)
    return out


def decoder_forward(params: dict, sampled_h, target, hidden: int = 512, num_layers: int = 2,
                    loss_type: str = "likelihood", drop_masks=None, drop_p: float = 0.0, torch_dropout: float = 0.0):
    """Decoder.forward (decoder.py:21-35); rnn dropout off unless keep masks are injected."""
    r = bilstm(params, sampled_h, hidden, num_layers, drop_masks, drop_p, torch_dropout)
    mean = fc_stack(params, "mean_fc.blocks", r)
    log_var = fc_stack(params, "log_var_fc.blocks", r)
    return {"mean": mean, "log_var": log_var, "rnn_out": r,
            "losses": {"recon_loss": recon_elementwise(mean, log_var, target, loss_type)}}


# ----------------------------------------------------------------------------
# length-masked reduction
# ----------------------------------------------------------------------------
def length_mask(lens: torch.Tensor, t_max: int) -> torch.Tensor:
    """mask[b, t] = arange(t_max)[t] < lens[b] * t_max evaluated in lens.dtype,
    with NO rounding (data_utils.py:88; length_to_mask builds
    arange(max_len, dtype=length.dtype) < length[:, None])."""
    scaled = lens * t_max
    return torch.arange(t_max, dtype=scaled.dtype)[None, :] < scaled[:, None]


def masked_reduce(loss: torch.Tensor, lens: torch.Tensor, reduction: str = "mean"):
    """apply_lens_to_loss (data_utils.py:67-104)."""
    m = length_mask(lens, loss.shape[1]).to(loss.dtype)
    while m.dim() < loss.dim():
        m = m.unsqueeze(-1)
    mask = torch.ones_like(loss) * m
    loss = loss * mask
    b = loss.shape[0]
    if reduction == "mean":
        return loss.sum() / mask.sum()
    if reduction == "batchmean":
        return loss.sum() / b
    if reduction == "batch":
        return loss.reshape(b, -1).sum(-1) / mask.reshape(b, -1).sum(-1)
    return loss


def weighted_total(losses: dict, hparams: dict):
    """compute_and_save_losses (md_model.py:189-213): x_loss -> x_weight (default 1);
    weights whose key contains '_kld' are divided by 2249 / batch_size."""
    total = 0
    for key, val in losses.items():
        wkey = key.replace("_loss", "_weight")
        w = hparams.get(wkey, 1)
        if "_kld" in wkey:
            w = w / (KLD_N_SAMPLES / hparams["batch_size"])
        total = total + w * val
    return total


def recipe_loss(enc_params: dict, dec_params: dict, feats, lens, eps, hparams: dict,
                hidden: int = 512, num_layers: int = 2, drop_masks=None, drop_p: float = 0.0, torch_dropout: float = 0.0):
    """test_vanilla_vae compute_forward + compute_objectives (model.py:19-55)
    after the normalizer: returns (loss, parts dict)."""
    enc = encoder_forward(enc_params, feats, eps)
    dec = decoder_forward(dec_params, enc["sampled_h"], feats, hidden, num_layers, drop_masks=drop_masks, drop_p=drop_p,
                          torch_dropout=torch_dropout)
    losses = {"kld_loss": masked_reduce(enc["loss"], lens),
              "recon_loss": masked_reduce(dec["losses"]["recon_loss"], lens)}
    return weighted_total(losses, hparams), {"enc": enc, "dec": dec, "losses": losses}


# ----------------------------------------------------------------------------
# InputNormalization(norm_type='global') [SB-recall]; declared at
# /root/reference/src/models/test_vanilla_vae/model.yaml:14-15, called at model.py:24-25
# ----------------------------------------------------------------------------
class GlobalNormRef:
    """Running global mean/std normaliser: per-utterance mean and unbiased std
    over round(len*T) valid frames, averaged over the batch, folded into a
    running average with weight 1/(count+1) while epoch < 3."""

    def __init__(self, update_until_epoch: int = 3, eps: float = 1e-10):
        self.count = 0
        self.glob_mean = None
        self.glob_std = None
        self.update_until_epoch = update_until_epoch
        self.eps = eps
        self.training = True

    def __call__(self, x, lens, epoch=0):
        means, stds = [], []
        for b in range(x.shape[0]):
            n = int(torch.round(lens[b] * x.shape[1]).int())
            seg = x[b, :n]
            means.append(seg.mean(0))
            stds.append(torch.clamp(seg.std(0), min=self.eps))
        cm, cs = torch.stack(means).mean(0), torch.stack(stds).mean(0)
        if self.training:
            if self.count == 0:
                self.glob_mean, self.glob_std = cm, cs
            elif epoch < self.update_until_epoch:
                w = 1 / (self.count + 1)
                self.glob_mean = (1 - w) * self.glob_mean + w * cm
                self.glob_std = (1 - w) * self.glob_std + w * cs
            self.count += 1
        return (x - self.glob_mean) / self.glob_std


def init_like_reference(input_size: int, enc_fc: int, latent: int, hidden: int, layers: int,
                        dec_fc: int, seed: int = 123456, dtype=torch.float32):
    """Random-init weights with the reference's key names and torch's default
    initialisers, in the reference's construction order (encoder then decoder,
    model.yaml:24-43) after torch.manual_seed(seed) (run.yaml:2-3)."""
    torch.manual_seed(seed)
    enc = {}
    for i, (a, b) in enumerate([(input_size, enc_fc), (enc_fc, enc_fc)]):
        lin = torch.nn.Linear(a, b)
        enc[f"fc.0.blocks.{2 * i}.weight"], enc[f"fc.0.blocks.{2 * i}.bias"] = lin.weight, lin.bias
    for name in ("mean_fc", "log_var_fc"):
        lin = torch.nn.Linear(enc_fc, latent)
        enc[f"{name}.weight"], enc[f"{name}.bias"] = lin.weight, lin.bias
    dec = {}
    lstm = torch.nn.LSTM(latent, hidden, layers, bidirectional=True, batch_first=True)
    for n, p in lstm.named_parameters():
        dec[f"rnn.{n}"] = p
    for name in ("mean_fc", "log_var_fc"):
        sizes = [2 * hidden, dec_fc, dec_fc, input_size]
        for i in range(3):
            lin = torch.nn.Linear(sizes[i], sizes[i + 1])
            dec[f"{name}.blocks.{2 * i}.weight"] = lin.weight
            dec[f"{name}.blocks.{2 * i}.bias"] = lin.bias
    f = lambda d: {k: v.detach().to(dtype).clone() for k, v in d.items()}
    return f(enc), f(dec)


def gaussian_nll_check():
    """log(2*pi) as float32, for documentation/tests."""
    return float(torch.log(2 * torch.tensor(np.pi))), math.log(2 * math.pi)
