"""CPU oracle: restatement of SpeechBrain 0.5.x ``lobes.features.Fbank``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the
arithmetic lives in the third-party ``speechbrain`` package, which is absent
from /root/reference and from this image, so this file restates the published
algorithm from the constants listed in SURVEY.md section 8c.

Reference call sites this follows:
  * declaration      /root/reference/src/config/run.yaml:39-44
      Fbank(deltas=True, sample_rate, hop_length[ms], n_fft, n_mels)
  * invocation       /root/reference/src/utils/data_io.py:197-201
      feat = compute_features(wav[None]).squeeze(0); drop the last frame when
      Kaldi (snip-edges=false) produced one frame fewer
  * Kaldi length     /root/reference/src/utils/data_io_utils.py:156
      compute-fbank-feats --snip-edges=false  ->  round-half-up(N / hop) frames

Pipeline restated (all SpeechBrain defaults unless the yaml overrides them):
  STFT      win = round(sr/1000 * 25 ms), hop = round(sr/1000 * hop_ms),
            torch.stft(n_fft, hop, win, hamming_window(win) [periodic],
            center=True, pad_mode='constant', onesided, not normalised)
  power     re^2 + im^2
  mel       n_mels triangular filters between f_min=0 and f_max=sr/2, the LEFT
            bandwidth of each filter used for both slopes
  dB        10*log10(clamp(x, 1e-10)), floor at (max over the utterance) - 80
  deltas    twice: 5-tap regression [-2..2]/10 with replicate padding along T
  output    cat([fbank, delta, delta-delta], -1)
"""
from __future__ import annotations

import math

import torch


def frame_counts(n_samples: int, hop: int) -> tuple[int, int]:
    """(frames SpeechBrain emits, frames kept after the Kaldi-length rule).

    data_io.py:198-201: the STFT (center=True) yields 1 + N // hop frames;
    Kaldi with --snip-edges=false yields floor(N / hop + 0.5); the reference
    asserts the difference is 0 or 1 and truncates to the Kaldi count.
    """
    t_sb = 1 + n_samples // hop
    t_kaldi = (n_samples + hop // 2) // hop
    t_keep = min(t_sb, t_kaldi)
    return t_sb, t_keep


def mel_filter_matrix(sample_rate: int = 16000, n_fft: int = 400, n_mels: int = 40,
                      f_min: float = 0.0, f_max: float | None = None,
                      dtype=torch.float32) -> torch.Tensor:
    """(n_fft//2+1, n_mels) triangular filter matrix, SpeechBrain Filterbank
    construction order (python-float mel end points, torch linspace/pow)."""
    if f_max is None:
        f_max = sample_rate / 2
    n_stft = n_fft // 2 + 1
    to_mel = lambda hz: 2595 * math.log10(1 + hz / 700)
    mel = torch.linspace(to_mel(f_min), to_mel(f_max), n_mels + 2, dtype=dtype)
    hz = 700 * (10 ** (mel / 2595) - 1)
    band = (hz[1:] - hz[:-1])[:-1]          # left bandwidth of every filter
    f_central = hz[1:-1]
    all_freqs = torch.linspace(0, sample_rate // 2, n_stft, dtype=dtype)
    slope = (all_freqs[None, :] - f_central[:, None]) / band[:, None]
    left, right = slope + 1.0, -slope + 1.0
    fb = torch.clamp(torch.minimum(left, right), min=0.0)      # (n_mels, n_stft)
    return fb.transpose(0, 1).contiguous()


def power_spectrum(wav: torch.Tensor, sample_rate=16000, hop_ms=10, win_ms=25,
                   n_fft=400) -> torch.Tensor:
    """(B, N) -> (B, 1+N//hop, n_fft//2+1) power spectrum."""
    win = int(round(sample_rate / 1000.0 * win_ms))
    hop = int(round(sample_rate / 1000.0 * hop_ms))
    window = torch.hamming_window(win, dtype=wav.dtype)
    spec = torch.stft(wav, n_fft, hop, win, window, center=True, pad_mode="constant",
                      normalized=False, onesided=True, return_complex=True)
    spec = torch.view_as_real(spec).transpose(2, 1)            # (B, T, F, 2)
    return spec.pow(2).sum(-1)


def amplitude_to_db(fbanks: torch.Tensor, amin=1e-10, top_db=80.0, multiplier=10.0,
                    ref_value=1.0) -> torch.Tensor:
    """(B, T, M) -> dB with the per-utterance top_db floor."""
    x_db = multiplier * torch.log10(torch.clamp(fbanks, min=amin))
    x_db = x_db - multiplier * math.log10(max(amin, ref_value))
    floor = x_db.amax(dim=(-2, -1)) - top_db
    return torch.maximum(x_db, floor.view(-1, 1, 1))


def deltas(x: torch.Tensor, window: int = 5) -> torch.Tensor:
    """(B, T, C) regression deltas, replicate padding along T."""
    n = (window - 1) // 2
    denom = n * (n + 1) * (2 * n + 1) / 3
    kernel = torch.arange(-n, n + 1, dtype=x.dtype)
    xt = x.transpose(1, 2)                                     # (B, C, T)
    xt = torch.nn.functional.pad(xt, (n, n), mode="replicate")
    c = xt.shape[1]
    out = torch.nn.functional.conv1d(xt, kernel.repeat(c, 1, 1), groups=c) / denom
    return out.transpose(1, 2)


def fbank(wav: torch.Tensor, deltas_: bool = True, sample_rate: int = 16000,
          hop_length: float = 10, n_fft: int = 400, n_mels: int = 40,
          dtype=torch.float32) -> torch.Tensor:
    """Oracle for ``Fbank(...)(wav)``: (B, N) -> (B, 1+N//hop, n_mels*(3|1)).

    The top_db floor is per row of ``wav`` (the reference only ever calls it
    with B == 1, data_io.py:197).
    """
    wav = wav.to(dtype)
    p = power_spectrum(wav, sample_rate, hop_length, 25, n_fft)
    fbm = mel_filter_matrix(sample_rate, n_fft, n_mels, dtype=dtype)
    fb = amplitude_to_db(torch.matmul(p, fbm))
    if not deltas_:
        return fb
    d1 = deltas(fb)
    d2 = deltas(d1)
    return torch.cat([fb, d1, d2], dim=2)


def audio_pipeline_features(wav_1d: torch.Tensor, deltas_: bool = True, sample_rate=16000,
                            hop_length=10, n_fft=400, n_mels=40,
                            dtype=torch.float32) -> torch.Tensor:
    """Oracle for what data_io.py:197-201 stores for one utterance: (N,) -> (T_keep, D)."""
    hop = int(round(sample_rate / 1000.0 * hop_length))
    feat = fbank(wav_1d[None], deltas_, sample_rate, hop_length, n_fft, n_mels, dtype)[0]
    t_sb, t_keep = frame_counts(wav_1d.shape[0], hop)
    assert feat.shape[0] == t_sb
    assert t_sb - t_keep in (0, 1)
    return feat[:t_keep]


def batched_features(wav: torch.Tensor, wav_lens: torch.Tensor, **kw):
    """Oracle for the batched front-end: per-utterance reference features,
    zero-padded to the longest kept length.  wav (B, N) zero padded, wav_lens
    (B,) int sample counts.  Returns (feat (B, T_max, D), frames (B,) int64)."""
    feats = [audio_pipeline_features(wav[b, : int(wav_lens[b])], **kw) for b in range(wav.shape[0])]
    t_max = max(f.shape[0] for f in feats)
    out = feats[0].new_zeros(len(feats), t_max, feats[0].shape[1])
    for b, f in enumerate(feats):
        out[b, : f.shape[0]] = f
    return out, torch.tensor([f.shape[0] for f in feats], dtype=torch.int64)
