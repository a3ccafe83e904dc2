"""GPU: the counter-based inter-layer dropout (csrc/dropout.cu) that stands in for nn.LSTM(dropout=p)
(modules/decoder.py:14-15, dec_rnn_dropout 0.15): mask bit-exact against the host restatement of the Philox stream,
backward = the same mask, and the whole decoder in training mode against the oracle with the same masks injected."""
import numpy as np
import pytest
import torch

from _util import FP32_RTOL, assert_close
from oracle import philox_ref, vae_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 7, 8, 1000, 4099, 1 << 20])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mask_is_bit_exact_and_scaled(cuda, n, dtype):
    from ml_vae_b200 import ops
    p, seed, off = 0.15, 0xABCDEF0123, 5
    x = (torch.randn(n, device=cuda) + 3.0).to(dtype)            # no zeros: y == 0 <=> dropped
    y = ops.dropout(x, p, seed, off)
    keep = torch.from_numpy(philox_ref.dropout_keep_mask(seed, off, n, p)).to(cuda)
    assert torch.equal(y != 0, keep)
    scale = torch.tensor(1.0, dtype=torch.float32) / (1.0 - torch.tensor(p, dtype=torch.float32))
    want = torch.where(keep, x.float() * scale.to(cuda), torch.zeros((), device=cuda)).to(dtype)
    assert torch.equal(y, want)
    # device-resident offset: offset + *counter
    ctr = torch.tensor([3], dtype=torch.int64, device=cuda)
    assert torch.equal(ops.dropout(x, p, seed, 2, ctr), y)
    if n >= 1000:
        assert abs(float(keep.float().mean()) - (1 - round(p * 65536) / 65536)) < 4 * (p * (1 - p) / n) ** 0.5


def test_backward_regenerates_the_mask(cuda):
    from ml_vae_b200 import ops
    x = torch.randn(33, 17, device=cuda, requires_grad=True)
    y = ops.dropout(x, 0.3, 9, 1)
    g = torch.randn_like(y)
    y.backward(g)
    keep = (y != 0).float()
    assert torch.equal(x.grad, g * keep * (torch.tensor(1.0) / (1.0 - torch.tensor(0.3))).to(cuda))
    assert ops.dropout(x, 0.0, 9, 1) is x


def test_decoder_training_mode_matches_oracle_with_injected_masks(cuda):
    """decoder.py:21-35 in train mode with rnn_dropout = 0.15 (fp32, 1e-5): forward values, loss and the gradients of
    both LSTM layers against the oracle decoder with the kernel's keep mask injected between the layers."""
    from ml_vae_b200.modules import Decoder
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(4)
    B, T, L, H, D, p = 3, 11, 8, 32, 12, 0.15
    dec = Decoder(L, H, 2, p, [2 * H, 16, 16, D]).to(cuda).train()
    dec.use_persistent_lstm = True
    z = torch.randn(B, T, L, device=cuda)
    tgt = torch.randn(B, T, D, device=cuda)
    lens = torch.tensor([1.0, 0.8, 0.5], device=cuda)

    # float32 goes through cuDNN layer by layer only when dropout must be ours: force the layered path
    out = dec(z, tgt, lens=lens)
    out["recon_loss"].backward()
    keep = philox_ref.dropout_keep_mask(dec.dropout_seed, 0, B * T * 2 * H, p).reshape(B, T, 2 * H)
    dp = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    ref = vae_ref.decoder_forward(dp, z.cpu(), tgt.cpu(), H, 2, drop_masks=[torch.from_numpy(keep)], drop_p=p)
    ref_loss = vae_ref.masked_reduce(ref["losses"]["recon_loss"], lens.cpu())
    ref_loss.backward()
    assert_close(out["mean"], ref["mean"], FP32_RTOL, "mean")
    assert_close(out["recon_loss"], ref_loss, FP32_RTOL, "loss")
    for k, prm in dec.named_parameters():
        assert_close(prm.grad, dp[k].grad, 5e-5, k)
    # eval mode: no dropout, the mask counter does not advance
    calls = dec.dropout_calls
    dec.eval()
    with torch.no_grad():
        ev = dec(z, tgt, lens=lens)
    ref0 = vae_ref.decoder_forward(dp, z.cpu(), tgt.cpu(), H, 2)
    assert_close(ev["mean"], ref0["mean"], FP32_RTOL, "eval mean")
    assert dec.dropout_calls == calls
