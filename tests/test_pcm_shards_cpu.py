"""CPU: the raw-PCM shard cache -- format round trip, ordering rules, data-parallel partition."""
import json
import os

import numpy as np
import pytest
import torch

from ml_vae_b200 import pcm_shards as ps


def _make(tmp_path, dtype="int16", n=23, shard_samples=5000):
    rng = np.random.default_rng(7)
    wavs = {}
    with ps.PcmShardWriter(str(tmp_path), 16000, dtype, shard_samples=shard_samples) as w:
        for i in range(n):
            ln = int(rng.integers(1, 3000))
            if dtype == "int16":
                x = rng.integers(-32768, 32768, ln).astype(np.int16)
                w.add(f"utt{i:03d}", x.astype(np.float32) / 32768.0, phn=np.arange(i % 5 + 1))   # float in, as librosa gives it
            else:
                x = rng.standard_normal(ln).astype(np.float32)
                w.add(f"utt{i:03d}", torch.from_numpy(x))
            wavs[f"utt{i:03d}"] = x
    return wavs


@pytest.mark.parametrize("dtype", ["int16", "float32"])
def test_round_trip_is_bit_exact(tmp_path, dtype):
    wavs = _make(tmp_path, dtype)
    r = ps.PcmShardReader(str(tmp_path))
    assert len(r) == len(wavs) and r.ids == sorted(wavs) and r.dtype == dtype
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".pcm")]) > 1          # rolled over to several shards
    for i, uid in enumerate(r.ids):
        assert r.num_samples(i) == wavs[uid].size
        assert np.array_equal(r.raw(i), wavs[uid])
        want = wavs[uid].astype(np.float32) / np.float32(32768.0) if dtype == "int16" else wavs[uid]
        assert np.array_equal(r.wav(i), want) and r.wav(i).dtype == np.float32           # librosa's float32 convention
        assert r.utts[i]["offset"] % ps.ALIGN == 0
    if dtype == "int16":
        assert np.array_equal(r.label(3, "phn"), np.arange(4))


def test_writer_rejects_bad_input(tmp_path):
    w = ps.PcmShardWriter(str(tmp_path))
    w.add("a", np.zeros(10, np.float32))
    with pytest.raises(ValueError):
        w.add("a", np.zeros(10, np.float32))              # duplicate id
    with pytest.raises(ValueError):
        w.add("b", np.full(4, 1.5, np.float32))           # outside [-1, 1) cannot be 16-bit PCM
    with pytest.raises(ValueError):
        w.add("c", np.zeros((2, 4), np.float32))          # not mono
    with pytest.raises(ValueError):
        w.add("d", np.zeros(0, np.float32))               # empty
    with pytest.raises(ValueError):
        ps.PcmShardWriter(str(tmp_path), dtype="int8")


def test_reader_detects_truncated_shard(tmp_path):
    _make(tmp_path)
    f = sorted(p for p in os.listdir(tmp_path) if p.endswith(".pcm"))[0]
    with open(tmp_path / f, "r+b") as fh:
        fh.truncate(os.path.getsize(tmp_path / f) - 16)
    with pytest.raises(ValueError):
        ps.PcmShardReader(str(tmp_path))
    idx = json.load(open(tmp_path / "index.json"))
    idx["version"] = 2
    json.dump(idx, open(tmp_path / "index.json", "w"))
    with pytest.raises(ValueError):
        ps.PcmShardReader(str(tmp_path))


def test_batch_order_sorting_and_partition():
    lengths = [5, 9, 1, 7, 7, 3, 8, 2, 6, 4, 10]
    desc = ps.batch_order(lengths, 4, "descending")
    assert [[lengths[i] for i in b] for b in desc] == [[10, 9, 8, 7], [7, 6, 5, 4], [3, 2, 1]]
    assert desc[0][3] == 3 and desc[1][0] == 4                                           # ties keep dataset order
    asc = ps.batch_order(lengths, 4, "ascending", drop_last=True)
    assert [[lengths[i] for i in b] for b in asc] == [[1, 2, 3, 4], [5, 6, 7, 7]]
    rnd = ps.batch_order(lengths, 4, "random", seed=1)
    assert sorted(i for b in rnd for i in b) == list(range(11)) and rnd == ps.batch_order(lengths, 4, "random", seed=1)
    with pytest.raises(ValueError):
        ps.batch_order(lengths, 4, "bogus")
    # data parallel: the ranks' slices of every global batch are disjoint, contiguous in the sorted order, and cover it
    for world in (2, 3):
        per_rank = [ps.batch_order(lengths, 2, "descending", world_size=world, rank=r) for r in range(world)]
        seen = [i for r in per_rank for b in r for i in b]
        assert len(seen) == len(set(seen))
        full = [i for b in ps.batch_order(lengths, 2 * world, "descending") for i in b]
        assert set(seen) <= set(full) and len(full) - len(seen) < world
        first = [per_rank[r][0] for r in range(world)]
        assert [i for b in first for i in b] == full[:2 * world]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_every_rank_sees_the_same_number_of_batches(world):
    """The step ends in an all-reduce: ranks with different batch counts would hang at the end of the epoch.
    Every remainder n % (B * world) is covered (the ceil split left the last ranks empty)."""
    B = 3
    for rem in range(0, B * world):
        n = 2 * B * world + rem
        lengths = [(7 * i) % 23 + 1 for i in range(n)]
        per_rank = [ps.batch_order(lengths, B, "descending", world_size=world, rank=r) for r in range(world)]
        counts = {len(p) for p in per_rank}
        assert len(counts) == 1, (world, rem, [len(p) for p in per_rank])
        assert all(len(b) >= 1 for p in per_rank for b in p)
        sizes = [len(p[-1]) for p in per_rank]
        assert max(sizes) - min(sizes) <= 1                                      # balanced trailing batch
        seen = [i for p in per_rank for b in p for i in b]
        assert len(seen) == len(set(seen))
        assert n - len(seen) == (rem if rem < world else 0)                      # only an un-shardable tail is dropped
