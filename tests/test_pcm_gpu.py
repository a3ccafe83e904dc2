"""GPU: ragged-PCM unpack kernel (bit-exact vs numpy) and the shard loader's on-device feature path:
features from the loader == the fused Fbank on the same padded float waveform, frame counts by the Kaldi rule."""
import numpy as np
import pytest
import torch

from ml_vae_b200 import _lib as L
from ml_vae_b200 import pcm_shards as ps

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["int16", "float32"])
def test_pcm_unpack_bit_exact(cuda, dtype):
    rng = np.random.default_rng(3)
    lens = [1, 7, 8, 9, 1000, 4099, 16000]
    npdt = np.int16 if dtype == "int16" else np.float32
    offs, pos, parts = [], 0, []
    for n in lens:
        x = rng.integers(-32768, 32768, n).astype(np.int16) if dtype == "int16" else rng.standard_normal(n).astype(np.float32)
        offs.append(pos)
        parts.append(np.concatenate([x, np.zeros((-n) % 8, npdt)]))
        pos += parts[-1].size
    blob = torch.from_numpy(np.concatenate(parts)).to(cuda)
    d_off = torch.tensor(offs, device=cuda)
    d_len = torch.tensor(lens, dtype=torch.int32, device=cuda)
    for n_max in (16000, 16001, 16004, 16384):
        out = torch.full((len(lens), n_max), float("nan"), device=cuda)
        scale = 1.0 / 32768.0 if dtype == "int16" else 1.0
        L.check(L.lib().mlvae_pcm_unpack(L.ptr(blob), 0 if dtype == "int16" else 1, L.ptr(d_off),
                                         L.ptr(d_len), len(lens), n_max, scale,
                                         L.ptr(out), L.stream_ptr()), "unpack")
        want = np.zeros((len(lens), n_max), np.float32)
        for b, (o, n) in enumerate(zip(offs, lens)):
            want[b, :n] = parts[b][:n].astype(np.float32) * np.float32(scale)
        assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("sorting", ["descending", "random"])
def test_loader_matches_fbank_on_padded_batch(cuda, tmp_path, sorting):
    from ml_vae_b200.features import Fbank
    rng = np.random.default_rng(11)
    wavs = {}
    with ps.PcmShardWriter(str(tmp_path), shard_samples=40000) as w:
        for i in range(13):
            n = int(rng.integers(800, 9000))
            x = np.clip(rng.standard_normal(n) * 3000, -32768, 32767).astype(np.int16)
            wavs[f"u{i}"] = x
            w.add(f"u{i}", x)
    reader = ps.PcmShardReader(str(tmp_path))
    fb = Fbank(deltas=True, hop_length=10, n_mels=40)
    loader = ps.PcmBatchLoader(reader, 4, fb, device=cuda, sorting=sorting)
    assert len(loader) == 4
    seen = []
    for batch in loader:
        ids = batch["id"]
        seen += ids
        feats, rel = batch["feat"]
        lens = [wavs[u].size for u in ids]
        n_max = max(lens)
        pad = torch.zeros(len(ids), n_max + (-n_max) % 4)
        for b, u in enumerate(ids):
            pad[b, :lens[b]] = torch.from_numpy(wavs[u].astype(np.float32) / np.float32(32768.0))
        want, want_rel = fb(pad.to(cuda), torch.tensor(lens, dtype=torch.int32, device=cuda), truncate=True)
        assert torch.equal(feats, want) and torch.equal(rel, want_rel)                 # same kernel, same float input: bit-exact
        wav, wrel = batch["wav"]
        assert torch.equal(wav.cpu(), pad)
        frames = [min(1 + n // 160, (n + 80) // 160) for n in lens]                    # data_io.py:199-201 (Kaldi frame count)
        assert feats.shape[1] == max(frames)
        assert torch.allclose(rel.cpu(), torch.tensor(frames, dtype=torch.float32) / max(frames))
        if sorting == "descending":
            assert lens == sorted(lens, reverse=True)
    assert sorted(seen) == sorted(wavs)
    ragged = sum(n + (-n) % 8 for n in (w.size for w in wavs.values())) * 2
    assert loader.h2d_bytes <= ragged + 16 * 13                                        # ragged int16, not padded float32
