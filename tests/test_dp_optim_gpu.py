"""Data-parallel optimiser step over peer memory (csrc/dp_optim.cu): W ranks simulated on ONE device.

Every "rank" has its own gradient / parameter / bf16 / sync arrays and its own stream; the ranks' kernels spin on each
other's epoch flags exactly as they do across NVLink, so the grids are capped (mlvae_dp_debug_max_ctas) to stay co-resident.
Checked against torch: g = sum_r grads_r / W, clip_grad_norm_, Adam (models/md_model.py:77-87 under DDP)."""
import ctypes as C

import pytest
import torch

from ml_vae_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference(params, grads_sum_scaled, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, max_norm=5.0):
    g = grads_sum_scaled.clone()
    norm = g.double().pow(2).sum().sqrt().float()
    coef = min(1.0, max_norm / (float(norm) + 1e-6)) if max_norm > 0 else 1.0
    g = g * coef
    m = m + (g - m) * (1 - b1)
    v = v * b2 + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p = params - (lr / bc1) * (m / (v.sqrt() / (bc2 ** 0.5) + eps))
    return p, m, v, float(norm)


class _Ranks:
    def __init__(self, W, n, seed=0):
        g = torch.Generator(device=DEV).manual_seed(seed)
        self.W, self.n = W, n
        p0 = torch.randn(n, device=DEV, generator=g)
        self.params = [p0.clone() for _ in range(W)]
        self.p16 = [p0.bfloat16() for _ in range(W)]
        self.grads = [torch.zeros(n, device=DEV) for _ in range(W)]
        self.m = [torch.zeros(n, device=DEV) for _ in range(W)]
        self.v = [torch.zeros(n, device=DEV) for _ in range(W)]
        self.sync = [torch.zeros(L.lib().mlvae_dp_sync_bytes(), dtype=torch.uint8, device=DEV) for _ in range(W)]
        self.loss = [torch.ones(1, device=DEV) for _ in range(W)]
        self.streams = [torch.cuda.Stream(DEV) for _ in range(W)]
        self.args = []
        for r in range(W):
            a = L.DpAdamArgs()
            a.world, a.rank, a.n = W, r, n
            for q in range(W):
                a.grads[q], a.params[q], a.params_bf16[q], a.sync[q] = (self.grads[q].data_ptr(), self.params[q].data_ptr(),
                                                                          self.p16[q].data_ptr(), self.sync[q].data_ptr())
            a.exp_avg, a.exp_avg_sq = self.m[r].data_ptr(), self.v[r].data_ptr()
            a.lr, a.beta1, a.beta2, a.eps, a.max_grad_norm = 1e-3, 0.9, 0.999, 1e-8, 5.0
            a.loss = self.loss[r].data_ptr()
            self.args.append(a)

    def step(self):
        torch.cuda.synchronize()
        for r in range(self.W):
            with torch.cuda.stream(self.streams[r]):
                L.check(L.lib().mlvae_dp_adam_step(self.args[r], C.c_void_p(self.streams[r].cuda_stream)), "mlvae_dp_adam_step", kernels=2)
        torch.cuda.synchronize()

    def state(self, r):
        out = (C.c_float * 5)()
        L.check(L.lib().mlvae_dp_read_state(L.ptr(self.sync[r]), C.byref(out), L.stream_ptr()), "read_state", kernels=0)
        return [float(x) for x in out]


@pytest.fixture(autouse=True)
def _cap_grids():
    L.lib().mlvae_dp_debug_max_ctas(24)
    yield
    L.lib().mlvae_dp_debug_max_ctas(0)


@pytest.mark.parametrize("W,n", [(1, 4096), (2, 100_000), (3, 65_544), (4, 1_000_000), (8, 262_144)])
def test_matches_torch_adam_over_steps(W, n):
    R = _Ranks(W, n)
    ref_p, ref_m, ref_v = R.params[0].clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(5)
    for step in range(1, 4):
        scale = 10.0 if step == 2 else 0.01                        # step 2 clips, the others do not
        gs = [scale * torch.randn(n, device=DEV, generator=gen) for _ in range(W)]
        for r in range(W):
            R.grads[r].copy_(gs[r])
        tot = gs[0].clone()
        for r in range(1, W):
            tot += gs[r]
        ref_p, ref_m, ref_v, norm = _reference(ref_p, tot / W, ref_m, ref_v, step)
        R.step()
        per = -(-(n // 4) // W) * 4
        for r in range(W):
            st = R.state(r)
            assert st[4] == 0.0, "a rank timed out waiting for its peers"
            assert st[0] == step and st[1] == step
            assert abs(st[2] - norm) <= 1e-5 * norm
            torch.testing.assert_close(R.params[r], ref_p, rtol=2e-6, atol=2e-7)
            assert torch.equal(R.p16[r], R.params[r].bfloat16())
            assert torch.equal(R.params[r], R.params[0])              # identical on every rank, bit for bit
            assert float(R.grads[r].abs().max()) == 0.0               # zero_grad over the whole local arena
            lo, hi = min(n, r * per), min(n, (r + 1) * per)
            torch.testing.assert_close(R.m[r][lo:hi], ref_m[lo:hi], rtol=2e-6, atol=1e-9)
            torch.testing.assert_close(R.v[r][lo:hi], ref_v[lo:hi], rtol=2e-6, atol=1e-12)


def test_non_finite_loss_on_one_rank_skips_everywhere():
    W, n = 3, 50_000
    R = _Ranks(W, n)
    before = R.params[0].clone()
    for r in range(W):
        R.grads[r].normal_()
    R.loss[1].fill_(float("nan"))
    R.step()
    for r in range(W):
        st = R.state(r)
        assert st[4] == 0.0 and st[0] == 1 and st[1] == 0             # epoch advanced, no optimiser step
        assert torch.equal(R.params[r], before)
        assert float(R.grads[r].abs().max()) == 0.0
    R.loss[1].fill_(1.0)
    for r in range(W):
        R.grads[r].normal_()
    R.step()
    assert all(R.state(r)[1] == 1 for r in range(W))
    assert not torch.equal(R.params[0], before)
    assert all(torch.equal(R.params[r], R.params[0]) for r in range(W))


def test_world_one_equals_single_gpu_step():
    n = 123_456 // 8 * 8
    R = _Ranks(1, n)
    p = R.params[0].clone(); p16 = p.bfloat16(); g = torch.randn(n, device=DEV) * 3
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    state = torch.zeros(L.lib().mlvae_adam_state_bytes() // 4, device=DEV)
    R.grads[0].copy_(g)
    L.check(L.lib().mlvae_adam_clip_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), L.ptr(p16), n, 1.0, 1e-3, 0.9, 0.999, 1e-8, 5.0, L.ptr(state), None,
                                         L.stream_ptr()), "adam", kernels=2)
    R.step()
    torch.testing.assert_close(R.params[0], p, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(R.m[0], m, rtol=1e-6, atol=1e-9)


def test_rejects_bad_arguments():
    a = L.DpAdamArgs()
    a.world, a.rank, a.n = 9, 0, 1024
    assert L.lib().mlvae_dp_adam_step(a, None) != 0
    a.world, a.rank, a.n = 2, 0, 1022
    assert L.lib().mlvae_dp_adam_step(a, None) != 0
