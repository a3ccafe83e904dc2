"""Raw-PCM shard cache and on-GPU batch path.

The reference computes features per utterance on the host (``audio_pipeline``, src/utils/data_io.py:189-202),
deep-copies every sample into one python dict and pickles it (data_io.py:67-97); training then loads the whole
pickle (data_io.py:101-137) and SpeechBrain pads batches on the host.  Here the cache holds what the features are a
pure function of -- the 16-bit PCM -- in flat shard files, and a batch travels to the GPU as ONE ragged blob
(half the bytes of the padded float32 batch or less); padding/scaling (``mlvae_pcm_unpack``) and the fused fbank
run on the device.  Batches come out ``PaddedBatch``-compatible: ``batch['feat'] -> (data, rel_lens)``,
``batch['wav'] -> (data, rel_lens)``, ``batch['id']``.

Layout of a cache directory:
    index.json          {"version": 1, "sample_rate", "dtype": "int16"|"float32", "align": 8,
                         "shards": [{"file", "num_samples"}], "utterances": [{"id", "shard", "offset", "num_samples"}]}
    shard-00000.pcm     little-endian samples, utterances back to back, every start aligned to ``align`` samples
    labels.npz          optional per-utterance integer/float label arrays, keys "<utt_id>/<name>"
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import _lib as L
from .brain import PaddedBatchLite
from .parallel import shard_bounds

ALIGN = 8                      # samples; keeps every utterance 16-byte aligned for the vector loads of the unpack kernel
_DTYPES = {"int16": np.dtype("<i2"), "float32": np.dtype("<f4")}


def _to_samples(wav, dtype: str) -> np.ndarray:
    a = wav.detach().cpu().numpy() if torch.is_tensor(wav) else np.asarray(wav)
    if a.ndim != 1:
        raise ValueError(f"expected a mono waveform (N,), got shape {a.shape}")
    if dtype == "int16":
        if a.dtype == np.int16:
            return a.astype("<i2", copy=False)
        # float waveform in [-1, 1) as librosa.load returns it for 16-bit files: x = s / 32768 exactly
        s = np.rint(a.astype(np.float64) * 32768.0)
        if np.any(s < -32768) or np.any(s > 32767):
            raise ValueError("waveform outside [-1, 1): store this cache as dtype='float32'")
        return s.astype("<i2")
    return a.astype("<f4", copy=False)


class PcmShardWriter:
    def __init__(self, directory: str, sample_rate: int = 16000, dtype: str = "int16", shard_samples: int = 1 << 27):
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        os.makedirs(directory, exist_ok=True)
        self.dir, self.sample_rate, self.dtype, self.shard_samples = directory, int(sample_rate), dtype, int(shard_samples)
        self.shards, self.utts, self.labels = [], [], {}
        self._f, self._pos, self._ids = None, 0, set()

    def _open_next(self):
        if self._f is not None:
            self._f.close()
            self.shards[-1]["num_samples"] = self._pos
        name = f"shard-{len(self.shards):05d}.pcm"
        self._f = open(os.path.join(self.dir, name), "wb")
        self.shards.append({"file": name, "num_samples": 0})
        self._pos = 0

    def add(self, utt_id: str, wav, **labels):
        if utt_id in self._ids:
            raise ValueError(f"duplicate utterance id {utt_id!r}")
        s = _to_samples(wav, self.dtype)
        if s.size == 0:
            raise ValueError(f"utterance {utt_id!r} is empty")
        if self._f is None or (self._pos > 0 and self._pos + s.size > self.shard_samples):
            self._open_next()
        self._f.write(s.tobytes())
        pad = (-s.size) % ALIGN
        if pad:
            self._f.write(np.zeros(pad, _DTYPES[self.dtype]).tobytes())
        self.utts.append({"id": utt_id, "shard": len(self.shards) - 1, "offset": self._pos, "num_samples": int(s.size)})
        self._pos += s.size + pad
        self._ids.add(utt_id)
        for k, v in labels.items():
            self.labels[f"{utt_id}/{k}"] = np.asarray(v)

    def close(self):
        if self._f is not None:
            self._f.close()
            self.shards[-1]["num_samples"] = self._pos
            self._f = None
        with open(os.path.join(self.dir, "index.json"), "w") as f:
            json.dump({"version": 1, "sample_rate": self.sample_rate, "dtype": self.dtype, "align": ALIGN,
                       "shards": self.shards, "utterances": self.utts}, f)
        if self.labels:
            np.savez(os.path.join(self.dir, "labels.npz"), **self.labels)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class PcmShardReader:
    def __init__(self, directory: str):
        with open(os.path.join(directory, "index.json")) as f:
            idx = json.load(f)
        if idx.get("version") != 1 or idx.get("align") != ALIGN or idx.get("dtype") not in _DTYPES:
            raise ValueError(f"{directory}: unsupported PCM cache (version/align/dtype)")
        self.dir, self.sample_rate, self.dtype = directory, int(idx["sample_rate"]), idx["dtype"]
        self.utts = idx["utterances"]
        self._maps = []
        for sh in idx["shards"]:
            path = os.path.join(directory, sh["file"])
            n = os.path.getsize(path) // _DTYPES[self.dtype].itemsize
            if n != sh["num_samples"]:
                raise ValueError(f"{path}: {n} samples on disk, index says {sh['num_samples']}")
            self._maps.append(np.memmap(path, dtype=_DTYPES[self.dtype], mode="r"))
        for u in self.utts:
            if u["offset"] % ALIGN or u["offset"] + u["num_samples"] > self._maps[u["shard"]].size:
                raise ValueError(f"{directory}: corrupt index entry for {u['id']!r}")
        lp = os.path.join(directory, "labels.npz")
        self._labels = np.load(lp) if os.path.exists(lp) else None

    def __len__(self):
        return len(self.utts)

    @property
    def ids(self):
        return [u["id"] for u in self.utts]

    def num_samples(self, i: int) -> int:
        return self.utts[i]["num_samples"]

    def raw(self, i: int) -> np.ndarray:
        u = self.utts[i]
        return self._maps[u["shard"]][u["offset"]:u["offset"] + u["num_samples"]]

    def wav(self, i: int) -> np.ndarray:
        """float32 waveform exactly as ``librosa.load`` yields it for a 16-bit file (data_io.py:193)."""
        r = self.raw(i)
        return r.astype(np.float32) * np.float32(1.0 / 32768.0) if self.dtype == "int16" else np.asarray(r)

    def label(self, i: int, name: str):
        return None if self._labels is None else self._labels[f"{self.utts[i]['id']}/{name}"]


def batch_order(lengths, batch_size: int, sorting: str = "descending", seed: int = 123456, drop_last: bool = False,
                world_size: int = 1, rank: int = 0):
    """Utterance indices of every batch THIS rank sees.  ``sorting`` as run.yaml:50 (ascending | descending | random,
    applied to the duration like the reference's sorted datasets).  Under data parallelism a global batch of
    ``batch_size * world_size`` utterances is cut into balanced contiguous per-rank slices (``parallel.shard_bounds``,
    SURVEY 8e); a trailing global batch that cannot give every rank at least one utterance is dropped, so every rank
    sees the SAME number of batches (the gradient all-reduce of the step would hang otherwise)."""
    n = len(lengths)
    if sorting == "ascending":
        order = sorted(range(n), key=lambda i: (lengths[i], i))
    elif sorting == "descending":
        order = sorted(range(n), key=lambda i: (-lengths[i], i))
    elif sorting == "random":
        order = np.random.default_rng(seed).permutation(n).tolist()
    else:
        raise ValueError(f"sorting must be ascending, descending or random, got {sorting!r}")
    gb = batch_size * world_size
    out = []
    for s in range(0, n, gb):
        chunk = order[s:s + gb]
        if len(chunk) < gb and (drop_last or len(chunk) < world_size):
            break
        lo, hi = shard_bounds(len(chunk), rank, world_size)      # >= 1 utterance per rank: len(chunk) >= world_size
        out.append(chunk[lo:hi])
    return out


class PcmBatchLoader:
    """Iterates ``PaddedBatchLite`` batches with the features computed on the GPU.

    Per batch: gather the utterances' raw samples into a pinned ragged blob (two staging sets, so the copy of batch
    i+1 overlaps the consumer's work on batch i), one H2D copy on a side stream, ``mlvae_pcm_unpack`` to the padded
    float32 matrix, then ``fbank(wav, wav_len, truncate=True)`` -> (feats, rel_lens).
    """

    def __init__(self, reader: PcmShardReader, batch_size: int, fbank, device="cuda:0", sorting: str = "descending",
                 seed: int = 123456, drop_last: bool = False, world_size: int = 1, rank: int = 0, out_dtype=torch.float32,
                 keep_wav: bool = True):
        self.reader, self.fbank, self.device = reader, fbank, torch.device(device)
        self.out_dtype, self.keep_wav = out_dtype, keep_wav
        lengths = [reader.num_samples(i) for i in range(len(reader))]
        self.batches = batch_order(lengths, batch_size, sorting, seed, drop_last, world_size, rank)
        self.h2d_bytes = 0
        self._stage = None
        aligned = [n + (-n) % ALIGN for n in lengths]
        self._cap = max((sum(aligned[i] for i in b) for b in self.batches), default=0)      # samples per staging blob
        self._max_b = max((len(b) for b in self.batches), default=0)

    def __len__(self):
        return len(self.batches)

    def _ensure_stage(self):
        if self._stage is None:
            tdt = torch.int16 if self.reader.dtype == "int16" else torch.float32
            self._stage = [{"blob": torch.empty(self._cap, dtype=tdt).pin_memory(),
                            "meta": torch.empty(2 * self._max_b, dtype=torch.int64).pin_memory(),
                            "dblob": torch.empty(self._cap, dtype=tdt, device=self.device),
                            "dmeta": torch.empty(2 * self._max_b, dtype=torch.int64, device=self.device),
                            "ready": torch.cuda.Event(), "free": torch.cuda.Event()} for _ in range(2)]
            for s in self._stage:
                s["free"].record()
            self._copy_stream = torch.cuda.Stream(self.device)

    def _submit(self, k: int, idxs):
        """Host gather + async H2D of one batch into staging set k."""
        lens = [self.reader.num_samples(i) for i in idxs]
        offs, pos = [], 0
        for n in lens:
            offs.append(pos)
            pos += n + (-n) % ALIGN
        self._ensure_stage()
        st = self._stage[k]
        st["free"].synchronize()                                  # the kernels that last read this set have finished
        blob = st["blob"].numpy()
        for i, o, n in zip(idxs, offs, lens):
            blob[o:o + n] = self.reader.raw(i)
        meta = st["meta"].numpy()
        meta[:len(idxs)] = offs
        meta[len(idxs):2 * len(idxs)] = lens
        with torch.cuda.stream(self._copy_stream):
            st["dblob"][:pos].copy_(st["blob"][:pos], non_blocking=True)
            st["dmeta"][:2 * len(idxs)].copy_(st["meta"][:2 * len(idxs)], non_blocking=True)
            st["ready"].record(self._copy_stream)
        self.h2d_bytes += pos * st["blob"].element_size() + 16 * len(idxs)
        return lens

    def _consume(self, k: int, idxs, lens):
        st = self._stage[k]
        B, n_max = len(idxs), max(lens)
        n_pad = n_max + (-n_max) % 4
        torch.cuda.current_stream(self.device).wait_event(st["ready"])
        wav = torch.empty(B, n_pad, dtype=torch.float32, device=self.device)
        d_off = st["dmeta"][:B]
        d_len = st["dmeta"][B:2 * B].to(torch.int32)
        scale = 1.0 / 32768.0 if self.reader.dtype == "int16" else 1.0
        L.check(L.lib().mlvae_pcm_unpack(L.ptr(st["dblob"]), 0 if self.reader.dtype == "int16" else 1, L.ptr(d_off), L.ptr(d_len),
                                         B, n_pad, scale, L.ptr(wav), L.stream_ptr()), "mlvae_pcm_unpack")
        st["free"].record()
        feats, rel = self.fbank(wav, d_len, truncate=True, out_dtype=self.out_dtype)
        batch = PaddedBatchLite({"id": [self.reader.utts[i]["id"] for i in idxs], "feat": (feats, rel), "wav_len": d_len})
        if self.keep_wav:
            batch["wav"] = (wav, d_len.float() / torch.full_like(d_len, n_max, dtype=torch.float32))   # exact IEEE division (see features.py)
        return batch

    def __iter__(self):
        if not self.batches:
            return
        pending = (0, self.batches[0], self._submit(0, self.batches[0]))
        for j in range(len(self.batches)):
            k, idxs, lens = pending
            if j + 1 < len(self.batches):
                pending = (k ^ 1, self.batches[j + 1], self._submit(k ^ 1, self.batches[j + 1]))
            yield self._consume(k, idxs, lens)
