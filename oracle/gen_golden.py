"""Generate tests/golden/*.npz from the REFERENCE ITSELF and validate the oracle.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

What it does
  1. puts oracle/sb_shim (stand-in for the absent `speechbrain`) and
     /root/reference/src on sys.path and imports, unmodified,
       modules.fc_block.FCBlock, modules.vanilla_vae.VanillaVAE,
       modules.decoder.Decoder, utils.data_utils.apply_lens_to_loss
  2. runs them (float32 and float64) on seeded inputs with eps injected by
     patching torch.randn_like, forward + backward, and stores inputs, weights,
     every output and every gradient as small .npz fixtures;
  3. asserts that oracle/vae_ref.py reproduces those outputs (so the oracle is
     PINNED for the VAE block);
  4. stores front-end fixtures from oracle/fbank_ref.py (PARITY UNPINNED: no
     SpeechBrain here) together with a float64 numpy DFT cross-check.

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = "/root/reference/src"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import fbank_np, fbank_ref, hvae_ref, philox_ref, vae_ref  # noqa: E402


def import_reference():
    if not os.path.isdir(REF_SRC):
        raise SystemExit("gen_golden.py needs /root/reference (build container only)")
    sys.path.insert(0, os.path.join(HERE, "sb_shim"))
    sys.path.insert(0, REF_SRC)
    from modules.vanilla_vae import VanillaVAE
    from modules.decoder import Decoder
    from utils.data_utils import apply_lens_to_loss
    return VanillaVAE, Decoder, apply_lens_to_loss


class inject_eps:
    """Make the reference's torch.randn_like(std) (vanilla_vae.py:39) return our eps."""

    def __init__(self, eps):
        self.eps = eps

    def __enter__(self):
        self._orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: self.eps.to(t.dtype).reshape(t.shape)

    def __exit__(self, *exc):
        torch.randn_like = self._orig


def np_dict(prefix, d):
    return {f"{prefix}{k}": v.detach().cpu().numpy() for k, v in d.items()}


def vae_case(name, B, T, D, L, enc_fc, hidden, layers, dec_fc, lens, seed, hp):
    VanillaVAE, Decoder, apply_lens_to_loss = import_reference()
    out = {}
    g = torch.Generator().manual_seed(seed)
    feats32 = torch.randn(B, T, D, generator=g)
    lens_t = torch.tensor(lens, dtype=torch.float32)
    eps_np = philox_ref.philox_normal(seed, 0, B * T * L)
    eps32 = torch.from_numpy(eps_np).reshape(B, T, L)
    out.update(feats=feats32.numpy(), lens=lens_t.numpy(), eps=eps_np.reshape(B, T, L),
               meta=np.array([B, T, D, L, enc_fc, hidden, layers, dec_fc, seed], np.int64),
               kld_weight=np.float64(hp.get("kld_weight", 1.0)), batch_size=np.int64(hp["batch_size"]))

    torch.manual_seed(seed)                       # run.yaml:2-3 precedes construction
    enc = VanillaVAE([D, enc_fc, enc_fc], L)      # model.yaml:24-29
    dec = Decoder(L, hidden, layers, 0.0, [2 * hidden, dec_fc, dec_fc, D])   # model.yaml:31-41, dropout 0
    out.update(np_dict("enc.", enc.state_dict()))
    out.update(np_dict("dec.", dec.state_dict()))

    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        e, d = enc.to(dt), dec.to(dt)
        e.zero_grad(); d.zero_grad()
        x = feats32.to(dt).clone().requires_grad_(True)
        with inject_eps(eps32), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            eo = e(x)
            do = d(eo["sampled_h"], x)
        kld = apply_lens_to_loss(eo["loss"], lens_t)
        rec = apply_lens_to_loss(do["losses"]["recon_loss"], lens_t)
        # md_model.py:189-213 (kld_weight has no '_kld' in it -> no rescale; recon weight defaults to 1)
        total = hp.get("kld_weight", 1.0) * kld + 1 * rec
        total.backward()
        res = {"mean": eo["mean"], "log_var": eo["log_var"], "sampled_h": eo["sampled_h"],
               "kld_elem": eo["loss"], "dec_mean": do["mean"], "dec_log_var": do["log_var"],
               "recon_elem": do["losses"]["recon_loss"], "kld_loss": kld, "recon_loss": rec,
               "total": total, "grad_feats": x.grad}
        out.update(np_dict(f"{tag}.", res))
        out.update({f"{tag}.grad.enc.{k}": p.grad.numpy().copy() for k, p in e.named_parameters()})
        out.update({f"{tag}.grad.dec.{k}": p.grad.numpy().copy() for k, p in d.named_parameters()})

        # ---- pin the oracle restatement against the reference ----
        ep = {k: v.detach().clone().requires_grad_(True) for k, v in e.state_dict().items()}
        dp = {k: v.detach().clone().requires_grad_(True) for k, v in d.state_dict().items()}
        x2 = feats32.to(dt).clone().requires_grad_(True)
        o_total, parts = vae_ref.recipe_loss(ep, dp, x2, lens_t,
                                             eps32.to(dt), {"kld_weight": hp.get("kld_weight", 1.0),
                                                            "batch_size": hp["batch_size"]}, hidden, layers)
        o_total.backward()
        tol = 1e-6 if dt == torch.float32 else 1e-12
        chk = [(o_total, total), (parts["enc"]["sampled_h"], eo["sampled_h"]), (parts["enc"]["loss"], eo["loss"]),
               (parts["dec"]["losses"]["recon_loss"], do["losses"]["recon_loss"]), (x2.grad, x.grad)]
        chk += [(ep[k].grad, p.grad) for k, p in e.named_parameters()]
        chk += [(dp[k].grad, p.grad) for k, p in d.named_parameters()]
        for a, b in chk:
            err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
            assert err <= tol, (name, tag, err)
    enc.float(); dec.float()
    np.savez_compressed(os.path.join(GOLD, f"vae_{name}.npz"), **out)
    print(f"vae_{name}: total f32={out['f32.total']:.7f} f64={out['f64.total']:.10f}  oracle==reference OK")


def mask_cases():
    """Frame predicate of apply_lens_to_loss for awkward (len, T) pairs, straight
    from the reference function, for all three reductions."""
    _, _, apply_lens_to_loss = import_reference()
    out = {}
    g = torch.Generator().manual_seed(7)
    for T in (7, 150, 300, 501, 2000):
        R = 16
        n = torch.randint(1, T + 1, (R,), generator=g)
        lens = (n.float() / T)
        lens[0] = 1.0
        loss = torch.randn(R, T, 2, generator=g)
        # per-row valid-frame count, from the reference ('batch' gives sum/count; use sum of mask via ones)
        valid = torch.stack([apply_lens_to_loss(torch.ones(1, T, 1), lens[b:b + 1], "batchmean") for b in range(R)])
        out[f"T{T}.lens"] = lens.numpy()
        out[f"T{T}.n_frames"] = n.numpy()
        out[f"T{T}.valid"] = valid.numpy().astype(np.int64)
        out[f"T{T}.loss"] = loss.numpy()
        for red in ("mean", "batchmean", "batch"):
            out[f"T{T}.{red}"] = apply_lens_to_loss(loss, lens, red).numpy()
    np.savez_compressed(os.path.join(GOLD, "mask_cases.npz"), **out)
    print("mask_cases written")


def fbank_cases():
    out = {}
    g = torch.Generator().manual_seed(123456)
    cases = [("a", 4000, 10, 80, False), ("b", 4321, 10, 80, True), ("c", 6400, 20, 40, True),
             ("d", 1600, 10, 40, True), ("e", 5119, 20, 40, False), ("f", 3360, 10, 80, True)]
    for tag, n, hop_ms, n_mels, dl in cases:
        wav = 0.1 * torch.randn(n, generator=g)
        if tag == "f":
            wav[1000:1800] = 0.0       # digital silence: exercises the 1e-10 clamp and the -80 dB floor
        kw = dict(deltas_=dl, hop_length=hop_ms, n_mels=n_mels)
        f32 = fbank_ref.audio_pipeline_features(wav, dtype=torch.float32, **kw)
        f64 = fbank_ref.audio_pipeline_features(wav, dtype=torch.float64, **kw)
        full64 = fbank_ref.fbank(wav[None], dl, 16000, hop_ms, 400, n_mels, torch.float64)[0]
        npv = fbank_np.fbank_np(wav.numpy(), dl, 16000, hop_ms, 400, n_mels)
        err = np.abs(npv - full64.numpy()).max()
        assert err < 1e-8, (tag, err)
        hop = 16 * hop_ms
        assert f32.shape[0] == min(1 + n // hop, (n + hop // 2) // hop)
        out[f"{tag}.wav"] = wav.numpy()
        out[f"{tag}.cfg"] = np.array([n, hop_ms, n_mels, int(dl)], np.int64)
        out[f"{tag}.f32"] = f32.numpy()
        out[f"{tag}.f64"] = f64.numpy()
        print(f"fbank_{tag}: N={n} hop={hop_ms}ms mels={n_mels} deltas={dl} -> {tuple(f32.shape)}; "
              f"|f32-f64|max={np.abs(f32.numpy() - f64.numpy()).max():.2e}; numpy-DFT vs torch.stft f64 {err:.1e}")
    np.savez_compressed(os.path.join(GOLD, "fbank_cases.npz"), **out)


def hvae_case():
    """GMM-VAE / hierarchical VAE from the reference's own modules, with the two randn_like draws
    (vanilla, then gmm: h_vae.py:31,38) and the Gumbel noise of F.gumbel_softmax (gmm_vae.py:31) injected."""
    import_reference()
    import torch.nn.functional as F
    from modules.h_vae import HierarchicalVAE
    B, T, D, L, N, fc = 3, 11, 24, 8, 3, 16
    g = torch.Generator().manual_seed(4242)
    feats = torch.randn(B, T, D, generator=g)
    pi_idx = torch.randint(0, 2, (B, T), generator=g).float()
    pi = torch.stack([1 - pi_idx, pi_idx], dim=2)
    eps_v = torch.randn(B, T, L, generator=g)
    eps_g = torch.randn(B, T, N * L, generator=g)
    gumbels = -torch.empty(B, T, N).exponential_(generator=g).log()
    torch.manual_seed(99)
    ref = HierarchicalVAE([D, fc, fc], L, N)
    out = {"feats": feats.numpy(), "pi": pi.numpy(), "eps_v": eps_v.numpy(), "eps_g": eps_g.numpy(), "gumbels": gumbels.numpy(),
           "meta": np.array([B, T, D, L, N, fc], np.int64)}
    out.update(np_dict("w.", ref.state_dict()))
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = ref.to(dt)
        m.zero_grad()
        x = feats.to(dt).clone().requires_grad_(True)
        queue = [eps_v.to(dt), eps_g.to(dt)]
        orig_randn, orig_gs = torch.randn_like, F.gumbel_softmax
        torch.randn_like = lambda t, *a, **k: queue.pop(0).reshape(t.shape)
        F.gumbel_softmax = lambda logits, tau=1, hard=False, **k: hvae_ref.gumbel_softmax_st(logits, gumbels.to(logits.dtype), tau)
        try:
            o = m(x, pi.to(dt))
        finally:
            torch.randn_like, F.gumbel_softmax = orig_randn, orig_gs
        cot = [torch.randn(B, T, L, generator=torch.Generator().manual_seed(7 + i)).to(dt) for i in range(4)]
        scalar = (o["mean"] * cot[0]).sum() + (o["log_var"] * cot[1]).sum() + (o["sampled_h"] * cot[2]).sum() \
            + (o["losses"]["vae_kld_loss"] * cot[3]).sum()
        scalar.backward()
        res = {"mean": o["mean"], "log_var": o["log_var"], "sampled_h": o["sampled_h"], "kld": o["losses"]["vae_kld_loss"],
               "gmm_weight": o["gmm_weight"], "grad_feats": x.grad}
        out.update(np_dict(f"{tag}.", res))
        out.update({f"{tag}.grad.{k}": p.grad.numpy().copy() for k, p in m.named_parameters()})
        if tag == "f32":
            out.update({f"cot{i}": c.float().numpy() for i, c in enumerate(cot)})
        # pin the restatement
        ps = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
        x2 = feats.to(dt).clone().requires_grad_(True)
        o2 = hvae_ref.hvae_forward(ps, x2, pi.to(dt), eps_v.to(dt), eps_g.to(dt), gumbels.to(dt))
        s2 = (o2["mean"] * cot[0]).sum() + (o2["log_var"] * cot[1]).sum() + (o2["sampled_h"] * cot[2]).sum() \
            + (o2["losses"]["vae_kld_loss"] * cot[3]).sum()
        s2.backward()
        tol = 2e-6 if dt == torch.float32 else 1e-12
        chk = [(o2["sampled_h"], o["sampled_h"]), (o2["losses"]["vae_kld_loss"], o["losses"]["vae_kld_loss"]), (x2.grad, x.grad)]
        chk += [(ps[k].grad, p.grad) for k, p in m.named_parameters()]
        for a, b in chk:
            err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
            assert err <= tol, ("hvae", tag, err)
    ref.float()
    np.savez_compressed(os.path.join(GOLD, "hvae_small.npz"), **out)
    print("hvae_small written; oracle==reference OK")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)           # deterministic summation order for the stored float32 values
    hp = {"kld_weight": 0.001, "batch_size": 8}
    # repo-default feature/latent sizes (run.yaml:26-29, model.yaml:18-23) with a small LSTM so weights fit a fixture
    vae_case("default_small", B=3, T=21, D=120, L=32, enc_fc=64, hidden=16, layers=2, dec_fc=64,
             lens=[1.0, 0.81, 0.33], seed=123456, hp=hp)
    # BASELINE config-1 sizes (80-dim fbank, latent 64)
    vae_case("c1_small", B=4, T=30, D=80, L=64, enc_fc=64, hidden=24, layers=2, dec_fc=64,
             lens=[1.0, 0.9, 0.5, 0.1], seed=20240, hp=hp)
    mask_cases()
    fbank_cases()
    hvae_case()


if __name__ == "__main__":
    main()
