"""GPU parity: the forward-only persistent LSTM layer (mlvae_lstm_fwd_dirs / _bwd_dirs with one direction) and the
``torch.nn.LSTM`` drop-in module built on it, against float32 ``torch.nn.LSTM`` on the same bf16-rounded parameters
(reference call sites: models/MD_VAE/model.yaml:78-83, modules/boundary_detector.py:19, modules/phoneme_recognizer.py:13).
Tolerance: BASELINE's bf16 bound, 1e-2 relative to the tensor's maximum."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-2


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def _ref(In, H, layers, bidir, dev, seed):
    torch.manual_seed(seed)
    ref = torch.nn.LSTM(In, H, layers, batch_first=True, bidirectional=bidir).to(dev)
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(p.bfloat16().float())
    return ref


@pytest.mark.parametrize("B,T,In,H", [(4, 6, 16, 32), (20, 33, 64, 128), (7, 40, 48, 256), (64, 500, 128, 512), (100, 60, 64, 512), (150, 20, 64, 512)])
def test_forward_only_layer_fwd_bwd_vs_torch(cuda, B, T, In, H):
    """(64, 500, 128, 512) is the MD_VAE recipe's main RNN at the benchmark batch; 100 rows run as ONE launch (112 CTAs),
    150 rows as two row chunks (144 + 6)."""
    from ml_vae_b200.lstm import lstm_layer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = _ref(In, H, 1, False, cuda, B + T + H)
    x = torch.randn(B, T, In, device=cuda).bfloat16()
    gy = torch.randn(B, T, H, device=cuda).bfloat16()
    xr = x.float().requires_grad_(True)
    yr, _ = ref(xr)
    (yr * gy.float()).sum().backward()
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    ps = [getattr(ref, n).detach().clone().requires_grad_(True) for n in names]
    xm = x.clone().requires_grad_(True)
    y = lstm_layer(xm, *ps, training=True)
    (y.float() * gy.float()).sum().backward()
    assert y.shape == (B, T, H) and y.dtype == torch.bfloat16
    assert rel(y, yr) < TOL, f"y {rel(y, yr):.2e}"
    assert rel(xm.grad, xr.grad) < TOL, f"dx {rel(xm.grad, xr.grad):.2e}"
    for n, p in zip(names, ps):
        assert rel(p.grad, getattr(ref, n).grad) < TOL, f"{n} {rel(p.grad, getattr(ref, n).grad):.2e}"
    # inference path (nothing saved) gives the same output, bit for bit
    with torch.no_grad():
        y2 = lstm_layer(x, *[p.detach() for p in ps], training=False)
    assert torch.equal(y, y2)


@pytest.mark.parametrize("bidir", [False, True])
def test_lstm_module_is_a_drop_in_for_nn_lstm(cuda, bidir):
    """Same constructor keywords, parameter names and return structure; a torch.nn.LSTM state_dict loads as it is."""
    from ml_vae_b200.modules import LSTM
    B, T, In, H, layers = 12, 50, 128, 512 if not bidir else 128, 2
    ref = _ref(In, H, layers, bidir, cuda, 3)
    m = LSTM(input_size=In, hidden_size=H, num_layers=layers, batch_first=True, dropout=0.0, bidirectional=bidir).to(cuda)
    assert [n for n, _ in m.named_parameters()] == [n for n, _ in ref.named_parameters()]
    m.load_state_dict(ref.state_dict())
    x = torch.randn(B, T, In, device=cuda).bfloat16()
    gy = torch.randn(B, T, H * (2 if bidir else 1), device=cuda).bfloat16()
    xr = x.float().requires_grad_(True)
    yr, (hr, _) = ref(xr)
    (yr * gy.float()).sum().backward()
    xm = x.clone().requires_grad_(True)
    out, (h_n, c_n) = m(xm)
    (out.float() * gy.float()).sum().backward()
    assert out.dtype == torch.bfloat16 and h_n.shape == hr.shape
    assert rel(out, yr) < TOL and rel(h_n, hr) < TOL and rel(xm.grad, xr.grad) < 2 * TOL
    for n, p in m.named_parameters():
        assert p.grad is not None and rel(p.grad, getattr(ref, n).grad) < 2 * TOL, n
    # float32 input: library validation path, same numbers as torch
    out32, (h32, c32) = m(x.float())
    assert torch.allclose(out32, ref(x.float())[0], atol=1e-5) and c32 is not None


def test_lstm_module_inter_layer_dropout_is_counter_based(cuda):
    """dropout=0.15 between the layers (MD_VAE/model.yaml:76): masks come from the (seed, call, layer) counter -- same seed and call
    give the same output, the next call a different one; eval() is deterministic and dropout-free."""
    from ml_vae_b200.modules import LSTM
    torch.manual_seed(0)
    x = torch.randn(8, 30, 64, device=cuda).bfloat16()
    a = LSTM(64, 128, 2, batch_first=True, dropout=0.15, seed=11).to(cuda)
    b = LSTM(64, 128, 2, batch_first=True, dropout=0.15, seed=11).to(cuda)
    b.load_state_dict(a.state_dict())
    a.train(); b.train()
    ya0, yb0 = a(x)[0], b(x)[0]
    ya1 = a(x)[0]
    assert torch.equal(ya0, yb0) and not torch.equal(ya0, ya1)
    a.eval(); b.eval()
    with torch.no_grad():
        e0, e1 = a(x)[0], a(x)[0]
    assert torch.equal(e0, e1) and not torch.equal(e0, ya0)
    xg = x.clone().requires_grad_(True)
    a.train()
    a(xg)[0].float().sum().backward()
    assert torch.isfinite(xg.grad.float()).all() and all(torch.isfinite(p.grad).all() for p in a.parameters())
