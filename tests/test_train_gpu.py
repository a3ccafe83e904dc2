"""GPU: the recipe (SBModel over the Brain-like loop) and the flat-arena TrainStep against one
full oracle step: features -> normaliser -> VAE -> losses -> backward -> clip -> Adam."""
import os

import numpy as np
import pytest
import torch

from _util import FP32_RTOL, assert_close
from conftest import GOLDEN
from oracle import fbank_ref, vae_ref

pytestmark = pytest.mark.gpu


def _oracle_step(enc_sd, dec_sd, x, lens, eps, hp, hidden, layers, lr=1e-3):
    ep = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in enc_sd.items()}
    dp = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in dec_sd.items()}
    loss, parts = vae_ref.recipe_loss(ep, dp, x, lens, eps, hp, hidden, layers)
    loss.backward()
    params = list(ep.values()) + list(dp.values())
    torch.nn.utils.clip_grad_norm_(params, 5.0)
    torch.optim.Adam(params, lr=lr).step()
    return loss.detach(), parts, ep, dp


def test_recipe_fit_batch_matches_oracle_step(cuda):
    from functools import partial
    from ml_vae_b200 import ops
    from ml_vae_b200.brain import EpochCounter, PaddedBatchLite
    from ml_vae_b200.models.b200_vanilla_vae.model import SBModel
    from ml_vae_b200.modules import Decoder, VanillaVAE
    from ml_vae_b200.normalizer import InputNormalization
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    z = np.load(os.path.join(GOLDEN, "vae_c1_small.npz"))
    B, T, D, L, enc_fc, hidden, layers, dec_fc, seed = [int(v) for v in z["meta"]]
    enc = VanillaVAE([D, enc_fc, enc_fc], L, seed=123456)
    dec = Decoder(L, hidden, layers, 0.0, [2 * hidden, dec_fc, dec_fc, D])
    enc.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("enc.")})
    dec.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("dec.")})
    enc_sd = {k: v.clone() for k, v in enc.state_dict().items()}
    dec_sd = {k: v.clone() for k, v in dec.state_dict().items()}
    hp = {"kld_weight": 0.001, "batch_size": 8, "optimizer": partial(torch.optim.Adam, lr=1e-3),
          "epoch_counter": EpochCounter(5), "normalizer": InputNormalization().to(cuda),
          "metric_keys": ["kld_loss", "recon_loss"]}
    model = SBModel(modules={"encoder": enc, "decoder": dec}, hparams=hp, run_opts={"device": "cuda:0"})
    model.on_stage_start("TRAIN")
    feats = torch.from_numpy(z["feats"]) * 3.0 + 1.5
    lens = torch.from_numpy(z["lens"])
    loss = model.fit_batch(PaddedBatchLite({"feat": (feats, lens)}))

    eps = ops.philox_normal((B, T, L), 123456, 0).cpu()           # what the kernel drew
    x = vae_ref.GlobalNormRef()(feats, lens)
    ref_loss, parts, ep, dp = _oracle_step(enc_sd, dec_sd, x, lens, eps, {"kld_weight": 0.001, "batch_size": 8}, hidden, layers)
    assert_close(loss, ref_loss, FP32_RTOL, "loss")
    assert_close(model.stats_loggers["kld_loss_stats"].loss_list[0], parts["losses"]["kld_loss"], FP32_RTOL, "kld")
    # Adam's first update is lr * g / (|g| + 1e-8): entries with |g| ~ 1e-8 amplify rounding, so the
    # updated weights are compared to within 5% of one lr step (gradient parity itself is checked at
    # 5e-5 in test_modules_gpu.py); at least 99.5% of the entries must agree to 1e-6.
    def upd(v, ref, what):
        d = (v.cpu() - ref.detach()).abs()
        assert float(d.max()) <= 0.05 * 1e-3, (what, float(d.max()))
        assert float((d <= 1e-6).float().mean()) > 0.995, what
    for k, v in enc.state_dict().items():
        upd(v, ep[k], f"updated enc.{k}")
    for k, v in dec.state_dict().items():
        upd(v, dp[k], f"updated dec.{k}")
    for p in enc.parameters():
        assert p.grad is None or float(p.grad.abs().sum()) == 0      # zero_grad ran


def test_train_step_from_audio_matches_oracle(cuda):
    from ml_vae_b200 import ops
    from ml_vae_b200.features import Fbank
    from ml_vae_b200.modules import Decoder, VanillaVAE
    from ml_vae_b200.normalizer import InputNormalization
    from ml_vae_b200.train_step import TrainStep
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(5)
    B, L, hidden, n_mels = 4, 16, 24, 40
    lens_abs = torch.tensor([6400, 6000, 4321, 1600])
    g = torch.Generator().manual_seed(8)
    wav = torch.zeros(B, 6400)
    for b in range(B):
        wav[b, : lens_abs[b]] = 0.1 * torch.randn(int(lens_abs[b]), generator=g)
    D = 3 * n_mels
    enc = VanillaVAE([D, 32, 32], L).to(cuda)
    dec = Decoder(L, hidden, 2, 0.0, [2 * hidden, 32, 32, D]).to(cuda)
    enc_sd = {k: v.clone() for k, v in enc.state_dict().items()}
    dec_sd = {k: v.clone() for k, v in dec.state_dict().items()}
    fb = Fbank(deltas=True, hop_length=10, n_mels=n_mels)
    ts = TrainStep(fb, InputNormalization().to(cuda), enc, dec, {"kld_weight": 0.001, "batch_size": B},
                   compute_dtype=torch.float32, seed=77)
    loss = ts.step(wav.to(cuda), lens_abs)

    feats, frames = fbank_ref.batched_features(wav, lens_abs, deltas_=True, hop_length=10, n_mels=n_mels)
    rel = frames.float() / feats.shape[1]
    x = vae_ref.GlobalNormRef()(feats, rel)
    eps = ops.philox_normal((B, feats.shape[1], L), 77, 0).cpu()
    ref_loss, parts, ep, dp = _oracle_step(enc_sd, dec_sd, x, rel, eps, {"kld_weight": 0.001, "batch_size": B}, hidden, 2)
    assert_close(loss, ref_loss, 5e-5, "loss")          # fbank (1e-5 on dB) -> normaliser -> model
    assert_close(ts.last["recon_loss"], parts["losses"]["recon_loss"], 5e-5, "recon")
    for k, v in enc.state_dict().items():
        d = (v.cpu() - ep[k].detach()).abs()
        assert float(d.max()) <= 0.1 * 1e-3 and float((d <= 5e-6).float().mean()) > 0.99, k
    assert float(ts.arena.grad.abs().sum()) == 0


def test_train_step_bf16_learns_and_skips_nonfinite(cuda):
    from ml_vae_b200.features import Fbank
    from ml_vae_b200.modules import Decoder, VanillaVAE
    from ml_vae_b200.normalizer import InputNormalization
    from ml_vae_b200.train_step import TrainStep
    torch.manual_seed(123456)
    B, n = 8, 16000
    wav = 0.1 * torch.randn(B, n, device=cuda)
    lens = torch.full((B,), n, dtype=torch.int32, device=cuda)
    enc = VanillaVAE([80, 64, 64], 64).to(cuda)
    dec = Decoder(64, 64, 2, 0.0, [128, 64, 64, 80]).to(cuda)
    ts = TrainStep(Fbank(deltas=False, hop_length=10, n_mels=80), InputNormalization().to(cuda), enc, dec,
                   {"kld_weight": 0.001, "batch_size": B}, compute_dtype=torch.bfloat16)
    losses = [float(ts.step(wav, lens)) for _ in range(30)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0] - 0.002
    before = ts.arena.flat.clone()
    bad = wav.clone()
    bad[0, 100] = float("inf")
    ts.step(bad, lens)                                   # non-finite loss: update skipped on the device
    assert not torch.isfinite(ts.last["loss"])
    assert torch.equal(ts.arena.flat, before)


def test_normalizer_kernel_matches_oracle(cuda):
    """InputNormalization(global) on csrc/norm.cu vs the oracle restatement, over several batches/epochs
    (first batch initialises, running average with weight 1/(count+1), frozen from epoch 3 on)."""
    from ml_vae_b200.normalizer import InputNormalization
    g = torch.Generator().manual_seed(9)
    ours, ref = InputNormalization().to(cuda), vae_ref.GlobalNormRef()
    for it, epoch in enumerate([0, 0, 1, 2, 3, 5]):
        B, T, D = 5, 37 + it, 24
        x = torch.randn(B, T, D, generator=g) * (1 + it) + it
        n = torch.randint(2, T + 1, (B,), generator=g)
        n[0] = T
        lens = n.float() / T
        want = ref(x, lens, epoch=epoch)
        got = ours(x.to(cuda), lens.to(cuda), epoch=epoch)
        assert_close(got, want, FP32_RTOL, f"batch {it}")
    assert ours.count == ref.count
    assert_close(ours.glob_std, ref.glob_std, FP32_RTOL, "running std")
    b16 = ours(x.to(cuda), lens.to(cuda), epoch=9, out_dtype=torch.bfloat16)
    assert b16.dtype == torch.bfloat16


def test_normalizer_statistics_exchange_equals_the_global_batch(cuda):
    """mlvae_global_norm_batch_avg -> all-reduce(sum) of the 2 D floats -> mlvae_global_norm_from_avg on every rank
    gives every rank the running statistics and the outputs of ONE normaliser that saw the global batch.  Two "ranks" are simulated
    on one device (the all-reduce is the sum of their two d_avg buffers)."""
    from ml_vae_b200 import _lib as L
    from ml_vae_b200.normalizer import InputNormalization
    lib = L.lib()
    g = torch.Generator().manual_seed(21)
    whole = InputNormalization().to(cuda)
    D, Bh = 24, 4
    states = [torch.zeros(lib.mlvae_norm_state_bytes(D) // 4, device=cuda) for _ in range(2)]
    for it, epoch in enumerate([0, 0, 1, 2, 4]):
        T = 31 + it
        x = (torch.randn(2 * Bh, T, D, generator=g) * (1 + it) + it).to(cuda)
        n = torch.randint(2, T + 1, (2 * Bh,), generator=g); n[0] = T
        lens = (n.float() / T).to(cuda)
        want = whole(x, lens, epoch=epoch)
        halves = [(x[r * Bh:(r + 1) * Bh].contiguous(), lens[r * Bh:(r + 1) * Bh].contiguous()) for r in range(2)]
        avgs, scratch = [], torch.empty(2 * Bh * D, device=cuda)
        for xr, lr in halves:
            a = torch.empty(2 * D, device=cuda)
            L.check(lib.mlvae_global_norm_batch_avg(L.ptr(xr), L.ptr(lr), Bh, T, D, L.ptr(scratch), L.ptr(a), L.stream_ptr()), "batch_avg")
            avgs.append(a)
        total = avgs[0] + avgs[1]                              # what all_reduce(SUM) leaves on every rank
        for r, (xr, lr) in enumerate(halves):
            out = torch.empty_like(xr)
            L.check(lib.mlvae_global_norm_from_avg(L.ptr(xr), Bh, T, D, L.ptr(total), 0.5, int(epoch < 3), L.ptr(states[r]), L.ptr(out),
                                                   L.F32, L.stream_ptr()), "from_avg")
            assert_close(out, want[r * Bh:(r + 1) * Bh], FP32_RTOL, f"batch {it} rank {r}")
    assert torch.equal(states[0], states[1])                   # identical running statistics on both ranks
    assert float(states[0][0]) == whole.count
    assert_close(states[0][4:4 + D], whole.glob_mean, FP32_RTOL, "running mean")
    assert_close(states[0][4 + D:4 + 2 * D], whole.glob_std, FP32_RTOL, "running std")
    # module switch: without an initialised process group it is the single-process normaliser
    solo = InputNormalization(sync_stats=True).to(cuda)
    ref2 = InputNormalization().to(cuda)
    assert torch.equal(solo(x, lens, epoch=0), ref2(x, lens, epoch=0))


def test_train_step_cuda_graph_replays_are_training_steps(cuda):
    """Captured step == eager step: same losses step by step (fresh eps each replay through the device counter,
    device-resident normaliser state, Adam)."""
    from ml_vae_b200.features import Fbank
    from ml_vae_b200.modules import Decoder, VanillaVAE
    from ml_vae_b200.normalizer import InputNormalization
    from ml_vae_b200.train_step import TrainStep

    def make():
        torch.manual_seed(11)
        enc = VanillaVAE([80, 64, 64], 64).to(cuda)
        dec = Decoder(64, 64, 2, 0.0, [128, 64, 64, 80]).to(cuda)
        return TrainStep(Fbank(deltas=False, hop_length=10, n_mels=80), InputNormalization().to(cuda), enc, dec,
                         {"kld_weight": 0.001, "batch_size": 8}, compute_dtype=torch.bfloat16, seed=5)

    g = torch.Generator().manual_seed(2)
    batches = [(0.1 * torch.randn(8, 8000, generator=g)).to(cuda) for _ in range(3)]
    lens = torch.full((8,), 8000, dtype=torch.int32, device=cuda)
    eager, graphed = make(), make()
    assert graphed.capture(batches[0], lens, warmup=2)
    for _ in range(2):                                                     # the capture ran 2 eager warm-up steps
        eager.step(batches[0], lens)
    seq_e, seq_g = [], []
    for i in range(6):
        seq_e.append(float(eager.step(batches[i % 3], lens)))
        seq_g.append(float(graphed.step(batches[i % 3], lens)))
    assert int(graphed.step_counter) == int(eager.step_counter)
    assert np.allclose(seq_e, seq_g, rtol=2e-2), (seq_e, seq_g)
    assert len(set(seq_g)) == len(seq_g)                                   # not a frozen replay


def test_train_step_resume_from_checkpoint_state(cuda):
    """Modules + normaliser + optimiser state + step counter restored into a fresh TrainStep: the next steps are bit-identical
    to the uninterrupted run (models/md_model.py:50-52: the reference registers the optimiser with its Checkpointer)."""
    from ml_vae_b200.features import Fbank
    from ml_vae_b200.modules import Decoder, VanillaVAE
    from ml_vae_b200.normalizer import InputNormalization
    from ml_vae_b200.train_step import TrainStep

    def make(seed):
        torch.manual_seed(seed)
        enc = VanillaVAE([80, 64, 64], 64).to(cuda)
        dec = Decoder(64, 64, 2, 0.15, [128, 64, 64, 80]).to(cuda)
        return TrainStep(Fbank(deltas=False, hop_length=10, n_mels=80), InputNormalization().to(cuda), enc, dec,
                         {"kld_weight": 0.001, "batch_size": 8}, compute_dtype=torch.bfloat16, seed=5)

    g = torch.Generator().manual_seed(3)
    batches = [(0.1 * torch.randn(8, 8000, generator=g)).to(cuda) for _ in range(3)]
    lens = torch.full((8,), 8000, dtype=torch.int32, device=cuda)
    a = make(1)
    for i in range(4):
        a.step(batches[i % 3], lens)
    ckpt = {"enc": {k: v.clone() for k, v in a.encoder.state_dict().items()}, "dec": {k: v.clone() for k, v in a.decoder.state_dict().items()},
            "norm": {k: (v.clone() if torch.is_tensor(v) else v) for k, v in a.normalizer.state_dict().items()},
            "opt": a.optimizer_state_dict(), "counter": a.step_counter.clone(), "calls": (a.encoder.calls, getattr(a.decoder, "dropout_calls", 0))}
    assert ckpt["opt"]["step"] == 4
    ref = [float(a.step(batches[i % 3], lens)) for i in range(4, 7)]
    b = make(99)                                                           # different initial weights: everything must come from the checkpoint
    b.encoder.load_state_dict(ckpt["enc"]); b.decoder.load_state_dict(ckpt["dec"]); b.normalizer.load_state_dict(ckpt["norm"])
    b.arena.refresh_bf16()
    b.load_optimizer_state_dict(ckpt["opt"])
    b.step_counter.copy_(ckpt["counter"])
    b.encoder.calls = ckpt["calls"][0]
    if hasattr(b.decoder, "dropout_calls"):
        b.decoder.dropout_calls = ckpt["calls"][1]
    got = [float(b.step(batches[i % 3], lens)) for i in range(4, 7)]
    assert got == ref, (got, ref)
    assert torch.equal(a.arena.flat, b.arena.flat)
