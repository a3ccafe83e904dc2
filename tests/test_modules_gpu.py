"""GPU parity of the drop-in modules (ml_vae_b200.modules) against golden vectors produced by
the reference's own VanillaVAE / Decoder / apply_lens_to_loss (tests/golden/vae_*.npz):
forward values, both losses, the weighted total and every parameter gradient."""
import os

import numpy as np
import pytest
import torch

from _util import BF16_RTOL, FP32_RTOL, assert_close
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _build(z, cuda, materialize=True):
    from ml_vae_b200.modules import Decoder, VanillaVAE
    B, T, D, L, enc_fc, hidden, layers, dec_fc, seed = [int(v) for v in z["meta"]]
    enc = VanillaVAE([D, enc_fc, enc_fc], L, materialize_loss=materialize)
    dec = Decoder(L, hidden, layers, 0.0, [2 * hidden, dec_fc, dec_fc, D], materialize_loss=materialize)
    enc.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("enc.")}, strict=True)
    dec.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("dec.")}, strict=True)
    return enc.to(cuda), dec.to(cuda)


@pytest.mark.parametrize("case", ["default_small", "c1_small"])
@pytest.mark.parametrize("fused", [False, True])
def test_modules_match_reference_golden_fp32(cuda, case, fused):
    from ml_vae_b200.utils.data_utils import apply_lens_to_loss
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    z = np.load(os.path.join(GOLDEN, f"vae_{case}.npz"))
    enc, dec = _build(z, cuda, materialize=not fused)
    feats = torch.from_numpy(z["feats"]).to(cuda).requires_grad_(True)
    lens = torch.from_numpy(z["lens"]).to(cuda)
    eps = torch.from_numpy(z["eps"]).to(cuda)
    kw = float(z["kld_weight"])
    if fused:       # masked means computed inside the loss kernels, nothing unreduced materialised
        eo = enc(feats, lens=lens, eps=eps)
        do = dec(eo["sampled_h"], feats, lens=lens)
        kld, rec = eo["kld_loss"], do["recon_loss"]
        assert eo["loss"] is None and do["losses"]["recon_loss"] is None
    else:           # the reference's contract: unreduced tensors + apply_lens_to_loss (model.py:47-50)
        eo = enc(feats, eps=eps)
        do = dec(eo["sampled_h"], feats)
        kld = apply_lens_to_loss(eo["loss"], lens)
        rec = apply_lens_to_loss(do["losses"]["recon_loss"], lens)
        assert_close(eo["loss"], torch.from_numpy(z["f64.kld_elem"]), FP32_RTOL, "kld_elem")
        assert_close(do["losses"]["recon_loss"], torch.from_numpy(z["f64.recon_elem"]), FP32_RTOL, "recon_elem")
    total = kw * kld + 1 * rec                         # md_model.py:189-213
    total.backward()
    g = lambda k: torch.from_numpy(np.asarray(z[f"f64.{k}"]))
    assert_close(eo["mean"], g("mean"), FP32_RTOL, "mean")
    assert_close(eo["log_var"], g("log_var"), FP32_RTOL, "log_var")
    assert_close(eo["sampled_h"], g("sampled_h"), FP32_RTOL, "sampled_h")
    assert_close(do["mean"], g("dec_mean"), FP32_RTOL, "dec_mean")
    assert_close(do["log_var"], g("dec_log_var"), FP32_RTOL, "dec_log_var")
    assert_close(kld, g("kld_loss"), FP32_RTOL, "kld_loss")
    assert_close(rec, g("recon_loss"), FP32_RTOL, "recon_loss")
    assert_close(total, g("total"), FP32_RTOL, "total")
    assert_close(feats.grad, g("grad_feats"), 5e-5, "grad_feats")
    for k, p in enc.named_parameters():
        assert_close(p.grad, g(f"grad.enc.{k}"), 5e-5, f"grad enc.{k}")
    for k, p in dec.named_parameters():
        assert_close(p.grad, g(f"grad.dec.{k}"), 5e-5, f"grad dec.{k}")


def test_modules_bf16_within_tolerance(cuda):
    z = np.load(os.path.join(GOLDEN, "vae_c1_small.npz"))
    enc, dec = _build(z, cuda, materialize=False)
    feats = torch.from_numpy(z["feats"]).to(cuda).bfloat16()
    lens = torch.from_numpy(z["lens"]).to(cuda)
    eps = torch.from_numpy(z["eps"]).to(cuda).bfloat16()
    eo = enc(feats, lens=lens, eps=eps)
    do = dec(eo["sampled_h"], feats, lens=lens)
    assert eo["sampled_h"].dtype == torch.bfloat16
    total = float(z["kld_weight"]) * eo["kld_loss"] + do["recon_loss"]
    total.backward()
    assert_close(total, torch.from_numpy(np.asarray(z["f64.total"])), BF16_RTOL, "bf16 total")
    assert_close(eo["kld_loss"], torch.from_numpy(np.asarray(z["f64.kld_loss"])), 3e-2, "bf16 kld")
    g = torch.from_numpy(z["f64.grad.dec.mean_fc.blocks.4.weight"])
    assert_close(dec.mean_fc.blocks._modules["4"].weight.grad, g, 5e-2, "bf16 grad")


@pytest.mark.parametrize("M,N", [(1, 8), (37, 80), (1000, 64), (32000, 128), (513, 1024)])
@pytest.mark.parametrize("leaky", [False, True])
def test_dense_bwd_prep_kernel(cuda, M, N, leaky):
    """g = dy * leaky'(y), db = column sums: the fused backward prologue of Linear(+LeakyReLU) (fc_block.py:9-16)."""
    from ml_vae_b200 import dense
    g0 = torch.Generator().manual_seed(M + N)
    dy = torch.randn(M, N, generator=g0).bfloat16().to(cuda)
    y = torch.randn(M, N, generator=g0).bfloat16().to(cuda)
    y[0, 0] = 0.0                                                  # LeakyReLU'(0) = slope (torch: y > 0 ? 1 : slope)
    g, db = dense._bwd_prep(dy, y if leaky else None)
    want_g = dy.float() * torch.where(y > 0, 1.0, 0.01).float() if leaky else dy.float()
    assert_close(g.float(), want_g, BF16_RTOL, "g")
    assert_close(db, g.double().sum(0), 1e-5, "db")               # exact column sums of the rounded g, fp32 accumulation
    g2, db2 = dense._bwd_prep(dy, y if leaky else None)           # self-resetting scratch, deterministic
    assert torch.equal(db, db2) and torch.equal(g, g2)
