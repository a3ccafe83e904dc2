"""Independent float64 numpy restatement of the front-end (explicit DFT matrix,
no torch.stft, no conv1d) used to bound the rounding of oracle/fbank_ref.py.

TEST INFRASTRUCTURE ONLY.  Pure numpy; O(T * 400 * 201) so keep inputs short.
Follows the same SpeechBrain 0.5.x constants as fbank_ref.py (PARITY UNPINNED).
"""
from __future__ import annotations

import numpy as np


def _mel_matrix(sample_rate, n_fft, n_mels):
    n_stft = n_fft // 2 + 1
    to_mel = lambda hz: 2595.0 * np.log10(1.0 + hz / 700.0)
    mel = np.linspace(to_mel(0.0), to_mel(sample_rate / 2.0), n_mels + 2)
    hz = 700.0 * (10.0 ** (mel / 2595.0) - 1.0)
    band = hz[1:-1] - hz[:-2]
    fc = hz[1:-1]
    freqs = np.linspace(0.0, sample_rate // 2, n_stft)
    slope = (freqs[:, None] - fc[None, :]) / band[None, :]
    return np.maximum(0.0, np.minimum(slope + 1.0, 1.0 - slope))   # (n_stft, n_mels)


def _delta(x):
    t = x.shape[0]
    idx = np.clip(np.arange(t)[:, None] + np.arange(-2, 3)[None, :], 0, t - 1)
    return (x[idx] * np.arange(-2, 3)[None, :, None]).sum(1) / 10.0


def fbank_np(wav_1d, deltas=True, sample_rate=16000, hop_ms=10, n_fft=400, n_mels=40):
    x = np.asarray(wav_1d, dtype=np.float64)
    win = int(round(sample_rate / 1000.0 * 25))
    hop = int(round(sample_rate / 1000.0 * hop_ms))
    assert win == n_fft
    n = x.shape[0]
    t = 1 + n // hop
    xp = np.concatenate([np.zeros(n_fft // 2), x, np.zeros(n_fft // 2 + hop)])
    frames = np.stack([xp[i * hop: i * hop + win] for i in range(t)])
    w = 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(win) / win)
    k = np.arange(n_fft // 2 + 1)
    ang = -2.0 * np.pi * np.outer(np.arange(win), k) / n_fft
    fw = frames * w
    power = (fw @ np.cos(ang)) ** 2 + (fw @ np.sin(ang)) ** 2
    fb = power @ _mel_matrix(sample_rate, n_fft, n_mels)
    db = 10.0 * np.log10(np.maximum(fb, 1e-10))
    db = np.maximum(db, db.max() - 80.0)
    if not deltas:
        return db
    d1 = _delta(db)
    return np.concatenate([db, d1, _delta(d1)], axis=1)
