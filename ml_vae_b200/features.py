"""Drop-in for ``speechbrain.lobes.features.Fbank`` as the reference declares it
(config/run.yaml:39-44) and calls it (utils/data_io.py:197-201), running as ONE
fused CUDA front-end (csrc/fbank.cu) instead of stft -> pow -> matmul -> log ->
clamp -> 2x grouped conv1d.

yaml change for a reference checkout:

    compute_features: !new:ml_vae_b200.features.Fbank      # was speechbrain.lobes.features.Fbank
        deltas: True
        sample_rate: !ref <sample_rate>
        hop_length: !ref <hop_length>
        n_fft: !ref <n_fft>
        n_mels: !ref <n_mels>

``Fbank(...)(wav)`` keeps SpeechBrain's contract: float32 (B, N) in, float32
(B, 1 + N // hop, n_mels * (3 if deltas else 1)) out, no gradient.  The batched
training-time variant ``forward(wav, wav_lens, truncate=True)`` additionally
applies the reference's Kaldi-length truncation (data_io.py:199-201) per
utterance and zero-fills past each utterance, returning (feats, rel_lens).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib as L


def speechbrain_mel_matrix(sample_rate: int, n_fft: int, n_mels: int) -> torch.Tensor:
    """(n_fft//2+1, n_mels) float32 triangular bank, SpeechBrain construction
    (f_min=0, f_max=sr/2, the left bandwidth of each filter on both slopes)."""
    top = 2595 * math.log10(1 + (sample_rate / 2) / 700)
    edges_hz = 700 * (10 ** (torch.linspace(0.0, top, n_mels + 2) / 2595) - 1)
    centre = edges_hz[1:-1]
    width = (edges_hz[1:] - edges_hz[:-1])[:-1]
    bins_hz = torch.linspace(0, sample_rate // 2, n_fft // 2 + 1)
    ramp = (bins_hz.unsqueeze(1) - centre.unsqueeze(0)) / width.unsqueeze(0)
    return torch.clamp(torch.minimum(ramp + 1.0, 1.0 - ramp), min=0.0).contiguous()


class Fbank(torch.nn.Module):
    def __init__(self, deltas=False, context=False, requires_grad=False, sample_rate=16000, f_min=0, f_max=None,
                 n_fft=400, n_mels=40, filter_shape="triangular", param_change_factor=1.0, param_rand_factor=0.0,
                 left_frames=5, right_frames=5, win_length=25, hop_length=10):
        super().__init__()
        if f_max is None:
            f_max = sample_rate / 2
        unsupported = []
        if context: unsupported.append("context=True")
        if requires_grad: unsupported.append("requires_grad=True")
        if filter_shape != "triangular": unsupported.append(f"filter_shape={filter_shape}")
        if f_min != 0 or f_max != sample_rate / 2: unsupported.append("f_min/f_max other than 0 / sr/2")
        if param_rand_factor != 0.0: unsupported.append("param_rand_factor")
        if unsupported:
            raise NotImplementedError("ml_vae_b200.features.Fbank implements the reference configuration only "
                                      f"(run.yaml:39-44); unsupported: {', '.join(unsupported)}")
        self.deltas = bool(deltas)
        self.sample_rate = int(sample_rate)
        self.n_fft = int(n_fft)
        self.n_mels = int(n_mels)
        self.hop = int(round(sample_rate / 1000.0 * hop_length))
        self.win = int(round(sample_rate / 1000.0 * win_length))
        if self.win != self.n_fft or self.n_fft != 400:
            raise NotImplementedError("only win_length == n_fft == 400 samples (25 ms @ 16 kHz, run.yaml:26-29)")
        self.feature_dim = self.n_mels * (3 if self.deltas else 1)
        self._plan = None
        self._plan_device = None
        self._scratch = None

    # -- plan / scratch management ------------------------------------------------------
    def _get_plan(self, device):
        if self._plan is None or self._plan_device != device:
            self._destroy()
            win = torch.hamming_window(self.win, dtype=torch.float32).contiguous()
            mel = speechbrain_mel_matrix(self.sample_rate, self.n_fft, self.n_mels)
            handle = C.c_void_p()
            with torch.cuda.device(device):
                L.check(L.lib().mlvae_fbank_plan_create(C.byref(handle), self.sample_rate, self.hop, self.n_fft,
                                                         self.n_mels, int(self.deltas), C.c_void_p(win.data_ptr()),
                                                         C.c_void_p(mel.data_ptr())), "mlvae_fbank_plan_create", kernels=0)
            self._plan, self._plan_device = handle, device
        return self._plan

    def _destroy(self):
        if self._plan is not None:
            L.lib().mlvae_fbank_plan_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def frames(self, n_samples: int, truncate: bool = False) -> int:
        full = 1 + n_samples // self.hop
        return min(full, (n_samples + self.hop // 2) // self.hop) if truncate else full

    # -- forward ------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, wav: torch.Tensor, wav_lens: torch.Tensor | None = None, truncate: bool = False,
                out_dtype: torch.dtype = torch.float32):
        """wav (B, N) float32 CUDA.  wav_lens: None, relative lengths (float, SpeechBrain
        convention, samples = round(rel * N)) or absolute sample counts (integer tensor)."""
        L.require_cuda(wav)
        if wav.dim() != 2:
            raise ValueError(f"expected (batch, time) waveform, got {tuple(wav.shape)}")
        wav = wav.float().contiguous()
        B, N = wav.shape
        plan = self._get_plan(wav.device)
        len_dev = None
        if wav_lens is not None:
            if wav_lens.dtype.is_floating_point:
                n_abs = torch.round(wav_lens.to(wav.device).float() * N).to(torch.int32)
            else:
                n_abs = wav_lens.to(device=wav.device, dtype=torch.int32)
            len_dev = torch.clamp(n_abs, 0, N).contiguous()
        t_out = self.frames(N, truncate)
        out = torch.empty(B, t_out, self.feature_dim, dtype=out_dtype, device=wav.device)
        frames = torch.empty(B, dtype=torch.int32, device=wav.device)
        need = L.lib().mlvae_fbank_scratch_bytes(plan, B, N)
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != wav.device:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=wav.device)
        L.check(L.lib().mlvae_fbank_fwd(plan, L.ptr(wav), L.ptr(len_dev), B, N, wav.stride(0), int(truncate),
                                        L.ptr(out), L.dtype_code(out), t_out, L.ptr(frames), L.ptr(self._scratch),
                                        L.stream_ptr()), "mlvae_fbank_fwd", kernels=2)
        if wav_lens is None and not truncate:
            return out
        # relative lengths exactly as the host pipeline forms them (IEEE float32 division): torch's CUDA division by a python
        # SCALAR multiplies by the reciprocal, which is 1 ulp off for some lengths and moves the reference's mask predicate
        # t < lens * T (data_utils.py:88) by one frame; tensor / tensor divides correctly rounded
        return out, frames.float() / torch.full((B,), float(t_out), dtype=torch.float32, device=wav.device)
