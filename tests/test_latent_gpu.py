"""GPU parity: fused reparam+KL, recon-loss and masked-reduce kernels (through the C ABI)
against the oracle (oracle/vae_ref.py) on identical inputs and identical eps.
Tolerances are the ones BASELINE.json states: fp32 1e-5 rel, bf16 1e-2 rel."""
import os

import numpy as np
import pytest
import torch

from _util import BF16_RTOL, FP32_RTOL, assert_close
from conftest import GOLDEN
from oracle import philox_ref, vae_ref

pytestmark = pytest.mark.gpu


def _lens(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    n = torch.randint(1, T + 1, (B,), generator=g)
    n[0] = T
    return (n.float() / T)


SHAPES = [(3, 21, 32), (4, 30, 64), (2, 7, 5), (1, 1, 1), (5, 33, 12), (8, 300, 64), (16, 129, 256), (3, 50, 1000)]


@pytest.mark.parametrize("B,T,L", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_reparam_kl_fwd_bwd_vs_oracle(cuda, B, T, L, dtype):
    from ml_vae_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + T * 10 + L)
    mu = torch.randn(B, T, L, generator=g).to(dtype)
    lv = (torch.randn(B, T, L, generator=g).clamp(-6, 3)).to(dtype)
    eps = torch.randn(B, T, L, generator=g).to(dtype)
    lens = _lens(B, T, 5)
    gz = torch.randn(B, T, L, generator=g).to(dtype)
    rtol = FP32_RTOL if dtype == torch.float32 else BF16_RTOL

    # oracle in float64 on the (possibly bf16-rounded) inputs; for fp32 also the fp32 oracle itself
    mu_r, lv_r = mu.double().requires_grad_(True), lv.double().requires_grad_(True)
    z_r = vae_ref.reparameterize(mu_r, lv_r, eps.double())
    kl_r = vae_ref.kld_elementwise(mu_r, lv_r)
    klm_r = vae_ref.masked_reduce(kl_r, lens)
    ((z_r * gz.double()).sum() + 0.37 * klm_r).backward()

    mu_d, lv_d = mu.to(cuda).requires_grad_(True), lv.to(cuda).requires_grad_(True)
    z, kl_elem, kl_mean = ops.reparam_kl(mu_d, lv_d, lens=lens.to(cuda), eps=eps.to(cuda), want_elem=True, want_mean=True)
    ((z.float() * gz.to(cuda).float()).sum() + 0.37 * kl_mean).backward()
    assert z.dtype == dtype and kl_mean.dtype == torch.float32
    assert_close(z.float(), z_r, rtol, "z")
    assert_close(kl_elem.float(), kl_r, rtol, "kl_elem")
    assert_close(kl_mean, klm_r, FP32_RTOL if dtype == torch.float32 else 2e-3, "kl_mean")
    assert_close(mu_d.grad.float(), mu_r.grad, rtol, "grad_mu")
    assert_close(lv_d.grad.float(), lv_r.grad, rtol, "grad_logvar")


@pytest.mark.parametrize("B,T,L", [(4, 30, 64), (2, 7, 5), (3, 50, 256), (2, 9, 12)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("philox", [False, True])
def test_reparam_kl_stacked_is_bit_identical_to_split(cuda, B, T, L, dtype, philox):
    """mlvae_reparam_kl_{fwd,bwd}_strided on [mean | log_var] (B, T, 2L) read in place (vanilla_vae.py:23-24 as ONE stacked
    projection) == the contiguous entry points on the two slices, bit for bit (same kernel, same arithmetic), incl. odd L
    (scalar path) and the Philox stream."""
    from ml_vae_b200 import ops
    g = torch.Generator().manual_seed(7 * B + T + L)
    ml = torch.randn(B, T, 2 * L, generator=g).clamp(-4, 3).to(dtype).to(cuda)
    eps = None if philox else torch.randn(B, T, L, generator=g).to(dtype).to(cuda)
    lens = _lens(B, T, 11).to(cuda)
    gz = torch.randn(B, T, L, generator=g).to(dtype).to(cuda)
    ge = torch.randn(B, T, L, generator=g).to(dtype).to(cuda)

    a = ml.clone().requires_grad_(True)
    z1, e1, m1 = ops.reparam_kl_stacked(a, lens=lens, eps=eps, seed=99, offset=3, want_elem=True, want_mean=True)
    ((z1.float() * gz.float()).sum() + (e1.float() * ge.float()).sum() + 0.37 * m1).backward()

    mu = ml[..., :L].clone().requires_grad_(True)
    lv = ml[..., L:].clone().requires_grad_(True)
    z2, e2, m2 = ops.reparam_kl(mu, lv, lens=lens, eps=eps, seed=99, offset=3, want_elem=True, want_mean=True)
    ((z2.float() * gz.float()).sum() + (e2.float() * ge.float()).sum() + 0.37 * m2).backward()

    assert torch.equal(z1, z2) and torch.equal(e1, e2) and torch.equal(m1, m2)
    assert a.grad.shape == ml.shape
    assert torch.equal(a.grad[..., :L], mu.grad) and torch.equal(a.grad[..., L:], lv.grad)


@pytest.mark.parametrize("B,T,L", [(4, 30, 64), (2, 7, 5)])
def test_reparam_kl_elem_gradient_path(cuda, B, T, L):
    """The reference module contract: unreduced 'loss' -> apply_lens_to_loss outside."""
    from ml_vae_b200 import ops
    from ml_vae_b200.utils.data_utils import apply_lens_to_loss
    g = torch.Generator().manual_seed(11)
    mu, lv, eps = (torch.randn(B, T, L, generator=g) for _ in range(3))
    lens = _lens(B, T, 6)
    mu_r, lv_r = mu.clone().requires_grad_(True), lv.clone().requires_grad_(True)
    (vae_ref.masked_reduce(vae_ref.kld_elementwise(mu_r, lv_r), lens, "batch").sum()
     + vae_ref.reparameterize(mu_r, lv_r, eps).sum()).backward()
    mu_d, lv_d = mu.to(cuda).requires_grad_(True), lv.to(cuda).requires_grad_(True)
    z, kl_elem, _ = ops.reparam_kl(mu_d, lv_d, eps=eps.to(cuda), want_elem=True, want_mean=False)
    (apply_lens_to_loss(kl_elem, lens.to(cuda), "batch").sum() + z.sum()).backward()
    assert_close(mu_d.grad, mu_r.grad, FP32_RTOL, "grad_mu")
    assert_close(lv_d.grad, lv_r.grad, FP32_RTOL, "grad_logvar")


def test_philox_stream_bit_exact_and_reproducible(cuda):
    from ml_vae_b200 import ops
    for seed, offset, n in [(123456, 0, 4099), (2 ** 40 + 17, 2 ** 33 + 5, 1000), (0, 0, 7)]:
        got = ops.philox_u32(n, seed, offset).cpu().numpy().view(np.uint32)
        assert np.array_equal(got, philox_ref.philox_u32(seed, offset, n))
        e = ops.philox_normal((n,), seed, offset).cpu().numpy()
        ref = philox_ref.philox_normal(seed, offset, n)
        assert np.abs(e - ref).max() < 5e-6
    # the fused kernel draws exactly the stream mlvae_philox_normal materialises (fwd and bwd)
    B, T, L = 4, 19, 64
    mu = torch.randn(B, T, L, device=cuda, requires_grad=True)
    lv = torch.randn(B, T, L, device=cuda, requires_grad=True)
    lens = torch.ones(B, device=cuda)
    z, _, klm = ops.reparam_kl(mu, lv, lens=lens, eps=None, seed=99, offset=3)
    eps = ops.philox_normal((B, T, L), 99, 3)
    z2, _, klm2 = ops.reparam_kl(mu, lv, lens=lens, eps=eps)
    assert torch.equal(z, z2) and torch.equal(klm, klm2)
    g1 = torch.autograd.grad(z.sum() + klm, [mu, lv])
    g2 = torch.autograd.grad(z2.sum() + klm2, [mu, lv])
    assert torch.equal(g1[0], g2[0]) and torch.equal(g1[1], g2[1])
    # different offset -> different eps
    z3, _, _ = ops.reparam_kl(mu, lv, lens=lens, eps=None, seed=99, offset=4)
    assert not torch.equal(z, z3)
    # the bf16 kernels draw their own stream (8 normals per Philox call from 16-bit uniforms, csrc/philox.cuh v2): device
    # materialisation == host restatement, and the fused bf16 kernel uses exactly that stream (forward and backward)
    for seed, offset, n in [(123456, 0, 4099), (2 ** 40 + 17, 2 ** 33 + 5, 1000), (0, 0, 7)]:
        e2 = ops.philox_normal((n,), seed, offset, kernel_dtype=torch.bfloat16).cpu().numpy()
        assert np.abs(e2 - philox_ref.philox_normal_v2(seed, offset, n)).max() < 5e-6
    eps2 = ops.philox_normal((B, T, L), 99, 3, kernel_dtype=torch.bfloat16)                 # float32 values of the bf16 stream
    assert not torch.equal(eps2, eps)
    mub, lvb = mu.detach().bfloat16().requires_grad_(True), lv.detach().bfloat16().requires_grad_(True)
    zb, _, klb = ops.reparam_kl(mub, lvb, lens=lens, seed=99, offset=3)
    zr = vae_ref.reparameterize(mub.detach().double(), lvb.detach().double(), eps2.double())
    assert_close(zb.float(), zr, BF16_RTOL, "bf16 philox z")
    gb = torch.autograd.grad(zb.float().sum() + klb, [mub, lvb])
    mur, lvr = mub.detach().double().cpu().requires_grad_(True), lvb.detach().double().cpu().requires_grad_(True)
    tot = vae_ref.reparameterize(mur, lvr, eps2.double().cpu()).sum() + vae_ref.masked_reduce(vae_ref.kld_elementwise(mur, lvr), lens.double().cpu())
    gr = torch.autograd.grad(tot, [mur, lvr])
    assert_close(gb[0].float(), gr[0], BF16_RTOL, "bf16 philox grad_mu")
    assert_close(gb[1].float(), gr[1], BF16_RTOL, "bf16 philox grad_logvar")          # 0.5 exp(0.5 lv) eps: carries the stream


@pytest.mark.parametrize("B,T,D", [(3, 21, 120), (4, 30, 80), (2, 9, 7), (8, 300, 240), (1, 1, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("loss_type", ["likelihood", "mse"])
def test_recon_fwd_bwd_vs_oracle(cuda, B, T, D, dtype, loss_type):
    from ml_vae_b200 import ops
    g = torch.Generator().manual_seed(B + T + D)
    mean = torch.randn(B, T, D, generator=g).to(dtype)
    lv = torch.randn(B, T, D, generator=g).clamp(-5, 3).to(dtype)
    tgt = torch.randn(B, T, D, generator=g).to(dtype)
    lens = _lens(B, T, 9)
    rtol = FP32_RTOL if dtype == torch.float32 else BF16_RTOL
    m_r, l_r, t_r = (x.double().requires_grad_(True) for x in (mean, lv, tgt))
    el_r = vae_ref.recon_elementwise(m_r, l_r, t_r, loss_type).double()
    red_r = vae_ref.masked_reduce(el_r, lens)
    (1.7 * red_r).backward()
    m_d, l_d, t_d = (x.to(cuda).requires_grad_(True) for x in (mean, lv, tgt))
    el, red = ops.recon_loss(m_d, l_d, t_d, lens=lens.to(cuda), loss_type=loss_type, want_elem=True, want_mean=True)
    (1.7 * red).backward()
    assert_close(el.float(), el_r, rtol, "elem")
    assert_close(red, red_r, FP32_RTOL if dtype == torch.float32 else 2e-3, "mean")
    assert_close(m_d.grad.float(), m_r.grad, rtol, "grad_mean")
    assert_close(t_d.grad.float(), t_r.grad, rtol, "grad_target")
    if loss_type == "likelihood":
        assert_close(l_d.grad.float(), l_r.grad, rtol, "grad_logvar")


def test_invalid_loss_type_raises(cuda):
    from ml_vae_b200 import ops
    x = torch.zeros(1, 2, 4, device=cuda)
    with pytest.raises(ValueError, match="Invalid loss type"):
        ops.recon_loss(x, x, x, lens=torch.ones(1, device=cuda), loss_type="huber")


def test_masked_reduce_matches_reference_golden(cuda):
    """Frame predicate bit-exact (valid-frame counts) and all three reductions, on the fixtures
    produced by the reference's own apply_lens_to_loss."""
    from ml_vae_b200.utils.data_utils import apply_lens_to_loss
    z = np.load(os.path.join(GOLDEN, "mask_cases.npz"))
    for T in (7, 150, 300, 501, 2000):
        lens = torch.from_numpy(z[f"T{T}.lens"]).to(cuda)
        R = lens.shape[0]
        ones = torch.ones(R, T, 1, device=cuda)
        counts = apply_lens_to_loss(ones, lens, "batchmean") * R          # sum(mask)
        assert int(round(float(counts))) == int(z[f"T{T}.valid"].sum())
        per_row = torch.stack([apply_lens_to_loss(ones[b:b + 1], lens[b:b + 1], "batchmean") for b in range(R)])
        assert np.array_equal(per_row.cpu().numpy().round().astype(np.int64).ravel(), z[f"T{T}.valid"].ravel())
        loss = torch.from_numpy(z[f"T{T}.loss"]).to(cuda)
        for red in ("mean", "batchmean", "batch"):
            got = apply_lens_to_loss(loss, lens, red).cpu()
            scale = vae_ref.masked_reduce(torch.from_numpy(z[f"T{T}.loss"]).abs(), lens.cpu(), red)
            assert ((got - torch.from_numpy(z[f"T{T}.{red}"])).abs() <= 2e-6 * scale + 1e-7).all(), (T, red)


@pytest.mark.parametrize("reduction", ["mean", "batchmean", "batch"])
def test_masked_reduce_backward(cuda, reduction):
    from ml_vae_b200.utils.data_utils import apply_lens_to_loss
    g = torch.Generator().manual_seed(3)
    loss = torch.randn(5, 37, 6, generator=g)
    lens = _lens(5, 37, 4)
    w = torch.randn(5 if reduction == "batch" else 1, generator=g)
    a = loss.clone().requires_grad_(True)
    (vae_ref.masked_reduce(a, lens, reduction) * w.squeeze()).sum().backward()
    b = loss.to(cuda).requires_grad_(True)
    (apply_lens_to_loss(b, lens.to(cuda), reduction) * w.to(cuda).squeeze()).sum().backward()
    assert_close(b.grad, a.grad, FP32_RTOL, "grad_loss")


def test_nonfinite_propagates_like_reference(cuda):
    """exp(logvar) overflow must give a non-finite loss exactly when the reference does
    (check_gradients then skips the step, md_model.py:82), including inf * 0 = NaN in padding."""
    from ml_vae_b200 import ops
    mu = torch.zeros(2, 4, 8)
    lv = torch.zeros(2, 4, 8)
    lv[1, 3, 2] = 100.0                     # in the masked-out tail of row 1
    lens = torch.tensor([1.0, 0.5])
    ref = vae_ref.masked_reduce(vae_ref.kld_elementwise(mu, lv), lens)
    _, _, got = ops.reparam_kl(mu.to(cuda), lv.to(cuda), lens=lens.to(cuda), eps=torch.zeros(2, 4, 8, device=cuda))
    assert torch.isnan(ref) and torch.isnan(got.cpu())
    lv[1, 3, 2] = 0.0
    lv[0, 0, 0] = 100.0
    ref = vae_ref.masked_reduce(vae_ref.kld_elementwise(mu, lv), lens)
    _, _, got = ops.reparam_kl(mu.to(cuda), lv.to(cuda), lens=lens.to(cuda), eps=torch.zeros(2, 4, 8, device=cuda))
    assert torch.isinf(ref) and torch.isinf(got.cpu())


def test_cpu_tensors_are_refused(lib_built):
    from ml_vae_b200 import ops
    from ml_vae_b200._lib import MlvaeError
    x = torch.zeros(1, 2, 4)
    with pytest.raises(MlvaeError, match="no CPU fallback"):
        ops.reparam_kl(x, x, lens=torch.ones(1))


def test_large_shape_properties(cuda):
    """BASELINE config-5 scale (16M frames x 64, bf16): size-independent properties.
    KL of (mu=0, logvar=0) is exactly 0; z == mu when eps == 0 ... and the masked mean of a
    constant is that constant regardless of lens."""
    from ml_vae_b200 import ops
    B, T, L = 64, 1 << 18, 64            # 16.8M frames, 1.07G elements
    mu = torch.zeros(B, T, L, device=cuda, dtype=torch.bfloat16)
    lv = torch.zeros_like(mu)
    lens = torch.linspace(0.3, 1.0, B, device=cuda)
    z, _, klm = ops.reparam_kl(mu, lv, lens=lens, seed=1, offset=0)
    assert float(klm) == 0.0
    # z = eps exactly (mu=0, std=1): moments of the Philox stream at scale
    s = z.float().mean().item(), z.float().pow(2).mean().item()
    assert abs(s[0]) < 1e-3 and abs(s[1] - 1) < 5e-3
    del z
    lv.fill_(1.0)
    mu.fill_(2.0)
    _, _, klm = ops.reparam_kl(mu, lv, lens=lens, seed=1, offset=0)
    want = -0.5 * (1 + 1.0 - 4.0 - np.exp(1.0))
    assert abs(float(klm) - want) < 1e-2 * abs(want)
