"""CPU: the oracle against the golden vectors produced by the reference's own modules
(tests/golden/*.npz, written by oracle/gen_golden.py in the build container)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fbank_np, fbank_ref, philox_ref, vae_ref


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def _params(z, prefix, dtype):
    return {k[len(prefix):]: torch.from_numpy(z[k]).to(dtype).requires_grad_(True)
            for k in z.files if k.startswith(prefix) and not k.startswith(("f32.", "f64."))}


@pytest.mark.parametrize("case", ["default_small", "c1_small"])
@pytest.mark.parametrize("tag,dtype,tol", [("f32", torch.float32, 2e-6), ("f64", torch.float64, 1e-12)])
def test_vae_oracle_matches_reference_golden(case, tag, dtype, tol):
    z = _load(f"vae_{case}.npz")
    B, T, D, L, enc_fc, hidden, layers, dec_fc, seed = [int(v) for v in z["meta"]]
    enc, dec = _params(z, "enc.", dtype), _params(z, "dec.", dtype)
    feats = torch.from_numpy(z["feats"]).to(dtype).requires_grad_(True)
    lens = torch.from_numpy(z["lens"])
    eps = torch.from_numpy(z["eps"]).to(dtype)
    hp = {"kld_weight": float(z["kld_weight"]), "batch_size": int(z["batch_size"])}
    total, parts = vae_ref.recipe_loss(enc, dec, feats, lens, eps, hp, hidden, layers)
    total.backward()

    def close(a, key):
        b = z[f"{tag}.{key}"]
        a = a.detach().numpy()
        assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30), key

    close(parts["enc"]["mean"], "mean")
    close(parts["enc"]["log_var"], "log_var")
    close(parts["enc"]["sampled_h"], "sampled_h")
    close(parts["enc"]["loss"], "kld_elem")
    close(parts["dec"]["mean"], "dec_mean")
    close(parts["dec"]["log_var"], "dec_log_var")
    close(parts["dec"]["losses"]["recon_loss"], "recon_elem")
    close(parts["losses"]["kld_loss"], "kld_loss")
    close(parts["losses"]["recon_loss"], "recon_loss")
    close(total, "total")
    close(feats.grad, "grad_feats")
    for k, p in enc.items():
        close(p.grad, f"grad.enc.{k}")
    for k, p in dec.items():
        close(p.grad, f"grad.dec.{k}")


def test_mask_predicate_matches_reference():
    z = _load("mask_cases.npz")
    for T in (7, 150, 300, 501, 2000):
        lens = torch.from_numpy(z[f"T{T}.lens"])
        m = vae_ref.length_mask(lens, T)
        assert np.array_equal(m.sum(1).numpy(), z[f"T{T}.valid"].ravel())       # bit-exact frame counts
        loss = torch.from_numpy(z[f"T{T}.loss"])
        for red in ("mean", "batchmean", "batch"):
            got = vae_ref.masked_reduce(loss, lens, red).numpy()
            # float32 summation order differs with the thread count: bound by the scale of the summands
            scale = np.abs(vae_ref.masked_reduce(loss.abs(), lens, red).numpy())
            assert (np.abs(got - z[f"T{T}.{red}"]) <= 2e-6 * scale + 1e-7).all(), (T, red)
    # the float32 product lens*T is NOT round(lens*T): the fixtures contain rows where they differ
    assert any((z[f"T{T}.valid"].ravel() != z[f"T{T}.n_frames"]).any() for T in (150, 300, 501, 2000))


def test_weighted_total_kld_rescale():
    # md_model.py:198-201: only weight KEYS containing '_kld' are rescaled by 2249 / batch_size
    losses = {"kld_loss": torch.tensor(2.0), "vae_kld_loss": torch.tensor(3.0), "recon_loss": torch.tensor(5.0)}
    hp = {"kld_weight": 0.5, "vae_kld_weight": 0.25, "batch_size": 8}
    want = 0.5 * 2.0 + (0.25 / (2249 / 8)) * 3.0 + 1 * 5.0
    assert abs(float(vae_ref.weighted_total(losses, hp)) - want) < 1e-6


def test_log2pi_is_float32_constant():
    f32, f64 = vae_ref.gaussian_nll_check()
    assert f32 == 1.8378770351409912          # the literal csrc/latent_loss.cu uses (kLog2Pi_f32)
    assert abs(f32 - f64) < 1e-6


def test_fbank_oracle_matches_golden_and_numpy_dft():
    z = _load("fbank_cases.npz")
    for tag in "abcdef":
        n, hop_ms, n_mels, dl = [int(v) for v in z[f"{tag}.cfg"]]
        wav = torch.from_numpy(z[f"{tag}.wav"])
        f32 = fbank_ref.audio_pipeline_features(wav, bool(dl), 16000, hop_ms, 400, n_mels, torch.float32)
        f64 = fbank_ref.audio_pipeline_features(wav, bool(dl), 16000, hop_ms, 400, n_mels, torch.float64)
        hop = 16 * hop_ms
        assert f32.shape == z[f"{tag}.f32"].shape
        assert f32.shape[0] == min(1 + n // hop, (n + hop // 2) // hop)          # frame count is integer exact
        assert np.allclose(f64.numpy(), z[f"{tag}.f64"], rtol=0, atol=1e-9)
        assert np.allclose(f32.numpy(), z[f"{tag}.f32"], rtol=0, atol=2e-3)       # float32 stft is build dependent
        ind = fbank_np.fbank_np(wav.numpy(), bool(dl), 16000, hop_ms, 400, n_mels)[: f64.shape[0]]
        assert np.abs(ind - f64.numpy()).max() < 1e-8


def test_frame_count_rule():
    # data_io.py:198-201: 1 + N//hop frames, drop one when Kaldi (round-half-up N/hop) has one fewer
    assert fbank_ref.frame_counts(48000, 160) == (301, 300)
    assert fbank_ref.frame_counts(80000, 160) == (501, 500)
    assert fbank_ref.frame_counts(320000, 160) == (2001, 2000)
    assert fbank_ref.frame_counts(48000, 320) == (151, 150)
    assert fbank_ref.frame_counts(4321, 160) == (28, 27)
    assert fbank_ref.frame_counts(4400, 160) == (28, 28)      # N % hop >= hop/2: nothing dropped


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, want in kat:
        got = philox_ref.philox4x32_10(np.array(c, np.uint32), np.array(k, np.uint32))
        assert [int(v) for v in got] == want


def test_philox_normal_v2_moments():
    """The bf16 kernels' stream (16-bit uniforms, 8 normals per Philox block): standard normal to 4 moments, |n| <= 4.71."""
    e = philox_ref.philox_normal_v2(123456, 0, 1 << 18).astype(np.float64)
    assert abs(e.mean()) < 6e-3 and abs(e.std() - 1) < 6e-3
    assert abs((e ** 3).mean()) < 3e-2 and abs((e ** 4).mean() - 3) < 6e-2
    assert np.abs(e).max() <= np.sqrt(-2 * np.log(2.0 ** -16)) + 1e-6
    assert not np.array_equal(e[:16], philox_ref.philox_normal(123456, 0, 16))


def test_philox_normal_moments():
    e = philox_ref.philox_normal(123456, 0, 1 << 18).astype(np.float64)
    assert abs(e.mean()) < 0.01 and abs(e.std() - 1) < 0.01
    assert abs((e ** 3).mean()) < 0.03 and abs((e ** 4).mean() - 3) < 0.1
    assert not np.array_equal(e[:16], philox_ref.philox_normal(123456, 1, 16))


@pytest.mark.parametrize("tag,dtype,tol", [("f32", torch.float32, 5e-6), ("f64", torch.float64, 1e-12)])
def test_hvae_oracle_matches_reference_golden(tag, dtype, tol):
    """oracle/hvae_ref.py against the reference's own HierarchicalVAE, forward and gradients."""
    from oracle import hvae_ref
    z = _load("hvae_small.npz")
    params = {k[2:]: torch.from_numpy(z[k]).to(dtype).requires_grad_(True) for k in z.files if k.startswith("w.")}
    t = lambda k: torch.from_numpy(z[k]).to(dtype)
    feats = t("feats").requires_grad_(True)
    o = hvae_ref.hvae_forward(params, feats, t("pi"), t("eps_v"), t("eps_g"), t("gumbels"))
    s = (o["mean"] * t("cot0")).sum() + (o["log_var"] * t("cot1")).sum() + (o["sampled_h"] * t("cot2")).sum() \
        + (o["losses"]["vae_kld_loss"] * t("cot3")).sum()
    s.backward()

    def close(a, key):
        b = z[f"{tag}.{key}"]
        a = a.detach().numpy()
        assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30), key

    close(o["mean"], "mean"); close(o["log_var"], "log_var"); close(o["sampled_h"], "sampled_h")
    close(o["losses"]["vae_kld_loss"], "kld"); close(o["gmm_weight"], "gmm_weight")
    close(feats.grad, "grad_feats")
    for k, p in params.items():
        close(p.grad, f"grad.{k}")


def test_front_end_pieces_agree_with_torchaudio():
    """Third-party anchor for the UNPINNED front-end oracle (SpeechBrain itself is absent): the three building blocks
    whose definitions coincide with torchaudio's -- the centred, zero-padded Hamming power STFT, the power-to-dB
    conversion with top_db = 80 and the 5-tap regression deltas with replicate padding -- are compared with
    torchaudio.transforms.Spectrogram / AmplitudeToDB / functional.compute_deltas.  The mel matrix is NOT compared:
    SpeechBrain's triangles (left bandwidth for both slopes) differ from torchaudio's by design."""
    torchaudio = pytest.importorskip("torchaudio")
    from oracle import fbank_ref
    g = torch.Generator().manual_seed(3)
    wav = 0.1 * torch.randn(2, 16000, generator=g, dtype=torch.float64)
    for hop_ms, hop in ((10, 160), (20, 320)):
        ours = fbank_ref.power_spectrum(wav, 16000, hop_ms, 25, 400)                       # (B, T, F)
        spec = torchaudio.transforms.Spectrogram(n_fft=400, win_length=400, hop_length=hop, window_fn=torch.hamming_window,
                                                 power=2.0, center=True, pad_mode="constant",
                                                 wkwargs={"dtype": torch.float64})(wav)    # (B, F, T)
        assert ours.shape == spec.transpose(1, 2).shape
        assert float((ours - spec.transpose(1, 2)).abs().max() / spec.abs().max()) < 1e-12
    mel = torch.rand(1, 50, 40, generator=g, dtype=torch.float64) * 10 ** (8 * torch.rand(1, 50, 40, generator=g, dtype=torch.float64) - 6)
    ours_db = fbank_ref.amplitude_to_db(mel)
    ta_db = torchaudio.transforms.AmplitudeToDB(stype="power", top_db=80.0)(mel)           # one "utterance": global max
    assert float((ours_db - ta_db).abs().max()) < 1e-10
    assert float(ours_db.min()) == pytest.approx(float(ours_db.max()) - 80.0)              # the floor is active in this case
    d_ours = fbank_ref.deltas(ours_db)
    d_ta = torchaudio.functional.compute_deltas(ours_db.transpose(1, 2), win_length=5, mode="replicate").transpose(1, 2)
    assert float((d_ours - d_ta).abs().max()) < 1e-12


def test_decode_oracle_matches_reference_golden():
    """oracle/decode_ref.py == the reference's decode_plvl_md_lbl_seqs_full (utils/decode_utils.py:374-565) on the vectors
    oracle/gen_golden_decode.py produced by running the reference itself: three integer sequences per utterance, bit-exact."""
    from oracle import decode_ref
    g = np.load(os.path.join(GOLDEN, "md_decode_cases.npz"))
    numpy2 = int(g["numpy_major"]) >= 2
    assert int(g["n_cases"]) >= 4
    for k in range(int(g["n_cases"])):
        c = lambda n: g[f"c{k}.{n}"]
        ob, of, op = decode_ref.decode_batch(c("log_p_yx"), c("log_p_b"), c("log_p_pi"), c("log_p_y"), c("y"), c("t_abs"), c("l_abs"),
                                             weight=float(c("weight")), numpy2=numpy2)
        for i in range(c("y").shape[0]):
            Ti, Li = int(c("t_abs")[i]), int(c("l_abs")[i])
            assert np.array_equal(ob[i], c("boundary")[i, :Ti])
            assert list(of[i]) == list(c("frames")[i, :Ti]) and list(op[i]) == list(c("phones")[i, :Li])
            assert int(ob[i].sum()) == Li and ob[i][0] == 1           # one boundary per phoneme, the first at frame 0
    # the reference's clamp-then-log helper (decode_utils.py:8-14) on its corner values
    x = np.array([0.0, 1e-6, 1e-5, 0.5, 1.0], dtype=np.float32)
    assert np.allclose(decode_ref.ref_log(x), np.log(np.array([1e-5, 1e-5, 1e-5, 0.5, 1.0], dtype=np.float32)))
    with pytest.raises(AssertionError):
        decode_ref.decode_one(np.zeros((2, 3, 2), np.float32), np.zeros((2, 2), np.float32), np.zeros((2, 2), np.float32),
                              np.zeros((3, 2), np.float32), np.array([0, 1, 2]))     # 3 phonemes, 2 frames
