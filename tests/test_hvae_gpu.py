"""GPU parity of the GMM-VAE / hierarchical-VAE family against golden vectors produced by the
reference's own HierarchicalVAE (tests/golden/hvae_small.npz): forward dict and every gradient, fp32 1e-5."""
import os

import numpy as np
import pytest
import torch

from _util import FP32_RTOL, assert_close
from conftest import GOLDEN
from oracle import hvae_ref

pytestmark = pytest.mark.gpu


def test_hvae_matches_reference_golden(cuda):
    from ml_vae_b200.modules import HierarchicalVAE
    torch.backends.cuda.matmul.allow_tf32 = False
    z = np.load(os.path.join(GOLDEN, "hvae_small.npz"))
    B, T, D, L, N, fc = [int(v) for v in z["meta"]]
    m = HierarchicalVAE([D, fc, fc], L, N)
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    assert sorted(m.state_dict()) == sorted(sd)                      # the reference's checkpoint keys
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda)
    t = lambda k: torch.from_numpy(z[k]).to(cuda)
    feats = t("feats").requires_grad_(True)
    o = m(feats, t("pi"), eps_vanilla=t("eps_v"), eps_gmm=t("eps_g"), gumbels=t("gumbels"))
    s = (o["mean"] * t("cot0")).sum() + (o["log_var"] * t("cot1")).sum() + (o["sampled_h"] * t("cot2")).sum() \
        + (o["losses"]["vae_kld_loss"] * t("cot3")).sum()
    s.backward()
    g = lambda k: torch.from_numpy(z[f"f64.{k}"])
    assert_close(o["mean"], g("mean"), FP32_RTOL, "mean")
    assert_close(o["log_var"], g("log_var"), FP32_RTOL, "log_var")
    assert_close(o["sampled_h"], g("sampled_h"), FP32_RTOL, "sampled_h")
    assert_close(o["losses"]["vae_kld_loss"], g("kld"), FP32_RTOL, "vae_kld_loss")
    assert torch.equal(o["gmm_weight"].cpu().round(), torch.from_numpy(z["f32.gmm_weight"]).round())   # one-hot choice: exact
    assert_close(feats.grad, g("grad_feats"), 5e-5, "grad_feats")
    for k, p in m.named_parameters():
        assert_close(p.grad, g(f"grad.{k}"), 5e-5, f"grad {k}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gmm_kernels_vs_oracle(cuda, dtype):
    from ml_vae_b200 import ops
    g = torch.Generator().manual_seed(5)
    shp = (3, 17, 3 * 8)
    mu, lv, pmu, plv, eps, gz, gk = (torch.randn(shp, generator=g).to(dtype) for _ in range(7))
    rtol = FP32_RTOL if dtype == torch.float32 else 1e-2
    r = [x.double().requires_grad_(True) for x in (mu, lv, pmu, plv)]
    zr = r[0] + torch.exp(0.5 * r[1]) * eps.double()
    kr = hvae_ref.gmm_kld_elementwise(r[2], r[3], r[0], r[1])
    ((zr * gz.double()).sum() + (kr * gk.double()).sum()).backward()
    d = [x.to(cuda).requires_grad_(True) for x in (mu, lv, pmu, plv)]
    zd, kd = ops.gmm_reparam_kl(d[0], d[1], d[2], d[3], eps=eps.to(cuda))
    ((zd.float() * gz.to(cuda).float()).sum() + (kd.float() * gk.to(cuda).float()).sum()).backward()
    assert_close(zd.float(), zr, rtol, "z")
    assert_close(kd.float(), kr, rtol, "kl")
    for a, b, n in zip(d, r, ("mu", "logvar", "prior_mu", "prior_logvar")):
        assert_close(a.grad.float(), b.grad, rtol, f"grad {n}")
    # apply_weight with general (not one-hot) weights, both gradients
    x = torch.randn(3, 17, 3, 8, generator=g).to(dtype)
    w = torch.rand(3, 17, 3, generator=g).to(dtype)
    go = torch.randn(3, 17, 8, generator=g).to(dtype)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    (hvae_ref.apply_weight(xr, wr) * go.double()).sum().backward()
    xd, wd = x.to(cuda).requires_grad_(True), w.to(cuda).requires_grad_(True)
    out = ops.apply_weight(xd, wd)
    (out.float() * go.to(cuda).float()).sum().backward()
    assert_close(out.float(), hvae_ref.apply_weight(x.double(), w.double()), rtol, "apply_weight")
    assert_close(xd.grad.float(), xr.grad, rtol, "grad x")
    assert_close(wd.grad.float(), wr.grad, rtol if dtype == torch.float32 else 3e-2, "grad w")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,N,C", [(5, 37, 3, 64), (2, 9, 4, 256), (3, 11, 2, 20), (1, 1, 5, 7)])
def test_apply_weight_shapes(cuda, dtype, B, T, N, C):
    """Vector path (C multiple of the 16-byte vector, power-of-two vectors per row) and the scalar fallback."""
    from ml_vae_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + C)
    x = torch.randn(B, T, N, C, generator=g).to(dtype)
    w = torch.rand(B, T, N, generator=g).to(dtype)
    go = torch.randn(B, T, C, generator=g).to(dtype)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    (hvae_ref.apply_weight(xr, wr) * go.double()).sum().backward()
    xd, wd = x.to(cuda).requires_grad_(True), w.to(cuda).requires_grad_(True)
    out = ops.apply_weight(xd, wd)
    (out.float() * go.to(cuda).float()).sum().backward()
    rtol = FP32_RTOL if dtype == torch.float32 else 1e-2
    assert_close(out.float(), hvae_ref.apply_weight(x.double(), w.double()), rtol, "apply_weight")
    assert_close(xd.grad.float(), xr.grad, rtol, "grad x")
    assert_close(wd.grad.float(), wr.grad, rtol if dtype == torch.float32 else 3e-2, "grad w")
