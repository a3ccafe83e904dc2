"""GPU: the TMA-fed tcgen05 GEMM (csrc/gemm.cu) against float32 torch matmuls of the same bf16 operands, for every operand
layout / epilogue the training step uses: x W^T (+bias, LeakyReLU), dA W (NN), dA^T x (both MN-major, float32 accumulate with
the LSTM gate-row permutation), the batched row-shifted dW_hh form, split-K, the fused dropout mask, grouped problems."""
import pytest
import torch

from _util import rel_err

pytestmark = pytest.mark.gpu


def _rand(shape, g, cuda, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).bfloat16().to(cuda)


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 0), (300, 64, 64, 64), (515, 520, 200, 0), (1000, 256, 1024, 256), (1000, 256, 1024, 128),
                                       (4001, 128, 1024, 0), (2048, 4096, 64, 0), (77, 8, 8, 0)])
@pytest.mark.parametrize("leaky", [False, True])
def test_nt_bias_activation(cuda, M, N, K, bn, leaky):
    """y = act(x W^T + b): both operands K-major (nn.Linear layout), bf16 out."""
    from ml_vae_b200.gemm import gemm
    g = torch.Generator().manual_seed(M + N + K)
    x, w = _rand((M, K), g, cuda), _rand((N, K), g, cuda, K ** -0.5)
    b = torch.randn(N, generator=g).to(cuda)
    y = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=cuda)
    gemm(x, w, y, M, N, K, lda=K, ldb=K, ldd=N, bias=b, leaky=leaky, bn=bn)
    ref = x.float() @ w.float().t() + b
    if leaky:
        ref = torch.nn.functional.leaky_relu(ref, 0.01)
    assert rel_err(y, ref) < 5e-3                       # bf16 rounding of the output only
    assert torch.isfinite(y.float()).all()


@pytest.mark.parametrize("M,N,K", [(256, 64, 4096), (1000, 1024, 512), (333, 72, 136)])
def test_nn_input_gradient_form(cuda, M, N, K):
    """dx = dA W with W row-major (K, N): A K-major, B MN-major."""
    from ml_vae_b200.gemm import gemm
    g = torch.Generator().manual_seed(M * 3 + N + K)
    a, w = _rand((M, K), g, cuda), _rand((K, N), g, cuda, K ** -0.5)
    d = torch.empty(M, N, dtype=torch.bfloat16, device=cuda)
    gemm(a, w, d, M, N, K, lda=K, ldb=N, ldd=N, b_mn=True)
    assert rel_err(d, a.float() @ w.float()) < 5e-3


@pytest.mark.parametrize("M,N,K,split", [(128, 64, 1000, 1), (512, 1024, 3000, 1), (2048, 64, 777, 1), (128, 1024, 32000, 8), (64, 80, 5000, 4)])
def test_tn_weight_gradient_form(cuda, M, N, K, split):
    """dW (+)= g^T x: both operands MN-major (row-major (K, M) and (K, N)), float32 out, accumulate, split-K."""
    from ml_vae_b200.gemm import gemm
    g = torch.Generator().manual_seed(M + 7 * N + K)
    a, x = _rand((K, M), g, cuda), _rand((K, N), g, cuda)
    d0 = torch.randn(M, N, generator=g).to(cuda)
    d = d0.clone()
    gemm(a, x, d, M, N, K, lda=M, ldb=N, ldd=N, a_mn=True, b_mn=True, out_f32=True, accumulate=True, split_k=split)
    ref = d0 + a.float().t() @ x.float()
    assert rel_err(d, ref) < 2e-5                       # float32 accumulation; only the summation order differs
    d2 = d0.clone()
    gemm(a, x, d2, M, N, K, lda=M, ldb=N, ldd=N, a_mn=True, b_mn=True, out_f32=True, accumulate=True, split_k=split)
    assert torch.equal(d, d2)                           # deterministic


def test_lstm_weight_gradient_forms(cuda):
    """The recurrent weight gradient as ONE batched GEMM over row-shifted views (no boundary corrections), both directions
    grouped in one launch, output rows permuted from the kernels' (unit, gate) order to torch's (gate, unit)."""
    from ml_vae_b200.gemm import gemm
    g = torch.Generator().manual_seed(5)
    Bb, T, H = 5, 37, 64
    dA = _rand((Bb, T, 2, 4 * H), g, cuda)              # (B, T, dir, unit*4+gate)
    y = _rand((Bb, T, 2 * H), g, cuda)
    out = [torch.zeros(4 * H, H, device=cuda) + 0.25 for _ in range(2)]
    dA2, y2 = dA.view(Bb * T, 8 * H), y.view(Bb * T, 2 * H)
    # forward direction: sum_b sum_{t>=1} dA[b,t,0]^T y[b,t-1,:H];  reverse: sum_b sum_{t<T-1} dA[b,t,1]^T y[b,t+1,H:]
    gemm([dA2[1:, : 4 * H], dA2[:, 4 * H:]], [y2[:, :H], y2[1:, H:]], out, 4 * H, H, T - 1, lda=8 * H, ldb=2 * H, ldd=H, a_mn=True, b_mn=True,
         kbatches=Bb, a_batch_stride=T * 8 * H, b_batch_stride=T * 2 * H, out_f32=True, accumulate=True, row_perm_H=H)
    inv = torch.arange(4 * H, device=cuda).view(H, 4).t().reshape(-1)          # torch row gate*H+unit <- kernel row unit*4+gate
    ref_f = torch.einsum("btr,bth->rh", dA[:, 1:, 0].float(), y[:, :-1, :H].float())[inv] + 0.25
    ref_r = torch.einsum("btr,bth->rh", dA[:, :-1, 1].float(), y[:, 1:, H:].float())[inv] + 0.25
    assert rel_err(out[0], ref_f) < 2e-5 and rel_err(out[1], ref_r) < 2e-5


def test_fused_dropout_epilogue_matches_the_dropout_kernel(cuda):
    from ml_vae_b200 import ops
    from ml_vae_b200.gemm import gemm
    g = torch.Generator().manual_seed(9)
    M, N, K = 700, 256, 512
    a, w = _rand((M, K), g, cuda), _rand((K, N), g, cuda, K ** -0.5)
    plain = torch.empty(M, N, dtype=torch.bfloat16, device=cuda)
    fused = torch.empty_like(plain)
    ctr = torch.tensor([4], dtype=torch.int64, device=cuda)
    gemm(a, w, plain, M, N, K, lda=K, ldb=N, ldd=N, b_mn=True)
    gemm(a, w, fused, M, N, K, lda=K, ldb=N, ldd=N, b_mn=True, drop_p=0.15, drop_seed=77, drop_offset=3, drop_offset_dev=ctr)
    want = ops.dropout(plain, 0.15, 77, 7)
    keep = want != 0
    assert torch.equal(fused != 0, keep) or float(((fused != 0) ^ keep).float().mean()) < 1e-4     # zeros of the plain product aside
    assert rel_err(fused, want) < 1e-2                  # mask applied before vs after the bf16 rounding
