"""GPU: the fused two-layer Linear / LeakyReLU chain kernels (csrc/mlp_chain.cu) against float32 torch on the same bf16
operands: forward values (hidden and output), input gradient, and the weight / bias gradients ACCUMULATED in place."""
import pytest
import torch

from _util import BF16_RTOL, rel_err

pytestmark = pytest.mark.gpu


def _views(N, K, g, cuda):
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    gw = torch.full((N, K), 0.5, device=cuda)                     # pre-existing gradient: must be accumulated into
    gb = torch.full((N,), -0.25, device=cuda)
    anchor = torch.zeros(1, device=cuda, requires_grad=True)
    return (w, b, gw, gb, anchor)


@pytest.mark.parametrize("M,KA,NA,NB,n,act_b", [(128, 64, 64, 80, 2, False), (300, 64, 64, 80, 2, False), (1000, 80, 64, 64, 1, True),
                                                 (4001, 80, 64, 64, 1, True), (32000, 64, 64, 80, 2, False), (77, 16, 16, 16, 1, False),
                                                 (513, 112, 112, 128, 1, True), (257, 48, 32, 96, 2, True)])
def test_chain2_fwd_bwd_vs_torch(cuda, M, KA, NA, NB, n, act_b):
    from ml_vae_b200.mlp_chain import chain2, supported
    assert supported(KA, NA, NB)
    g = torch.Generator().manual_seed(M + KA + NA + NB)
    x = torch.randn(M, n * KA, generator=g).bfloat16().to(cuda).requires_grad_(True)
    va = [_views(NA, KA, g, cuda) for _ in range(n)]
    vb = [_views(NB, NA, g, cuda) for _ in range(n)]
    gy = [torch.randn(M, NB, generator=g).bfloat16().to(cuda) for _ in range(n)]
    outs = chain2(x, va, vb, act_b)
    torch.autograd.backward(outs, gy)
    xr = x.detach().float().requires_grad_(True)
    for i in range(n):
        wa, ba, wb, bb = (t.detach().float().requires_grad_(True) for t in (va[i][0], va[i][1], vb[i][0], vb[i][1]))
        h = torch.nn.functional.leaky_relu(xr[:, i * KA:(i + 1) * KA] @ wa.t() + ba, 0.01)
        hq = h.bfloat16().float()                                  # the kernel feeds the bf16-rounded hidden activation on
        y = (h + (hq - h).detach()) @ wb.t() + bb
        if act_b:
            y = torch.nn.functional.leaky_relu(y, 0.01)
        assert rel_err(outs[i], y) < BF16_RTOL, ("y", i)
        y.backward(gy[i].float(), retain_graph=True)
        assert rel_err(va[i][2] - 0.5, wa.grad) < BF16_RTOL, ("dW_a", i)
        assert rel_err(va[i][3] + 0.25, ba.grad) < BF16_RTOL, ("db_a", i)
        assert rel_err(vb[i][2] - 0.5, wb.grad) < BF16_RTOL, ("dW_b", i)
        assert rel_err(vb[i][3] + 0.25, bb.grad) < BF16_RTOL, ("db_b", i)
    assert rel_err(x.grad, xr.grad) < BF16_RTOL
    # deterministic
    x2 = x.detach().clone().requires_grad_(True)
    for v in va + vb:
        v[2].fill_(0.5); v[3].fill_(-0.25)
    first = [t.clone() for v in va + vb for t in (v[2], v[3])]
    for v in va + vb:
        v[2].fill_(0.5); v[3].fill_(-0.25)
    torch.autograd.backward(chain2(x2, va, vb, act_b), gy)
    # (first was taken after the fill, so compare two fresh runs instead)
    second = [t.clone() for v in va + vb for t in (v[2], v[3])]
    for v in va + vb:
        v[2].fill_(0.5); v[3].fill_(-0.25)
    x3 = x.detach().clone().requires_grad_(True)
    torch.autograd.backward(chain2(x3, va, vb, act_b), gy)
    third = [t.clone() for v in va + vb for t in (v[2], v[3])]
    assert all(torch.equal(a, b) for a, b in zip(second, third)) and torch.equal(x2.grad, x3.grad)
