"""Drop-in for the reference's utils/decode_utils.py::decode_plvl_md_lbl_seqs_full (:374-565; the MD_VAE* recipes import it as
``decode_plvl_md_lbl_seqs``, e.g. models/MD_VAE/model.py:20): same arguments, same three returned lists, integer outputs equal
to the reference's bit for bit.

The reference runs the dynamic programme as a python triple loop per utterance under joblib (seconds per batch); here the
whole batch is one launch of ``mlvae_md_decode`` (csrc/md_decode.cu, one CTA per utterance).  The pre-computation
(decode_utils.py:417-438: sigmoid / softmax / stack on the model's device, then the clamped ``log`` ON THE CPU, :8-14) is kept
call for call, so the kernel sees exactly the float32 log-probabilities the reference's loop sees; ``device_log=True`` does the
clamp + log on the GPU instead (no host round trip; the logs may differ from the CPU's in the last ulp, so a label can flip at
an exact near-tie).  No CPU fallback: the decoder itself only exists as the CUDA kernel.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L

_EPS = 1e-5


def log(x: torch.Tensor, device_log: bool = False) -> torch.Tensor:
    """decode_utils.py:8-14 (returns a tensor instead of a numpy array)."""
    ret = x.detach().clone() if device_log else x.detach().cpu().clone()
    ret[torch.logical_and(ret >= 0, ret < _EPS)] = _EPS
    return torch.log(ret)


def _numpy_is_v2() -> bool:
    return int(np.__version__.split(".")[0]) >= 2


def decode_from_logs(log_p_yx, log_p_b, log_p_pi, log_p_y, y, feat_lens_abs, seq_lens_abs, weight=1.0, numpy2=None, device=None):
    """The kernel on ready-made log-probabilities: log_p_yx (B, T, N, 2), log_p_b (B, T, 2), log_p_pi (B, T, 2), log_p_y (N, 2)
    float32; y (B, Lmax) int; absolute lengths (B).  -> int32 tensors boundary (B, T), frames (B, T), phones (B, Lmax), status (B)."""
    dev = torch.device(device) if device is not None else (log_p_yx.device if log_p_yx.is_cuda else torch.device("cuda"))
    f = lambda t: torch.as_tensor(t).to(device=dev, dtype=torch.float32).contiguous()
    i32 = lambda t: torch.as_tensor(t).to(device=dev, dtype=torch.int32).contiguous()
    log_p_yx, log_p_b, log_p_pi, log_p_y = f(log_p_yx), f(log_p_b), f(log_p_pi), f(log_p_y)
    y, tl, sl = i32(y), i32(feat_lens_abs), i32(seq_lens_abs)
    B, T, N, two = log_p_yx.shape
    if two != 2 or log_p_b.shape != (B, T, 2) or log_p_pi.shape != (B, T, 2) or log_p_y.shape != (N, 2) or y.shape[0] != B:
        raise ValueError("decode_from_logs: inconsistent shapes")
    Lmax = y.shape[1]
    lib = L.lib()
    need = lib.mlvae_md_decode_workspace_bytes(B, T, Lmax)
    ws = torch.empty(need, dtype=torch.uint8, device=dev) if need else None
    boundary = torch.empty(B, T, dtype=torch.int32, device=dev)
    frames = torch.empty(B, T, dtype=torch.int32, device=dev)
    phones = torch.empty(B, Lmax, dtype=torch.int32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.mlvae_md_decode(L.ptr(log_p_yx), L.ptr(log_p_b), L.ptr(log_p_pi), L.ptr(log_p_y), L.ptr(y), L.ptr(tl), L.ptr(sl),
                                    B, T, N, Lmax, float(weight), int(_numpy_is_v2() if numpy2 is None else numpy2), L.ptr(ws),
                                    L.ptr(boundary), L.ptr(frames), L.ptr(phones), L.ptr(status), L.stream_ptr()), "mlvae_md_decode")
    return boundary, frames, phones, status


def decode_plvl_md_lbl_seqs_full(predictions, utt_ids, feat_lens, plvl_cnnl_seqs, plvl_cnnl_seq_lens, prior, weight=1.0,
                                 device_log: bool = False):
    """Decode the boundaries and the pi sequence simultaneously (decode_utils.py:374-565).

    predictions: {'phn_recog_out' (B, T, N) logits, 'boundary_v' (B, T), 'pi_logits' (B, T, 2)}; feat_lens / plvl_cnnl_seq_lens
    relative lengths (B); plvl_cnnl_seqs (B, L) canonical phoneme indices; prior (N).
    -> (decoded_boundary_seqs: list of int arrays (T_i,), flvl_md_lbl_seqs: list of lists (T_i), plvl_md_lbl_seqs: list of lists (L_i))
    """
    out = predictions["phn_recog_out"]
    L.require_cuda(out)
    dev = out.device
    # absolute lengths (decode_utils.py:413-414)
    t_abs = torch.round(feat_lens * out.shape[1]).int()
    l_abs = torch.round(plvl_cnnl_seq_lens * plvl_cnnl_seqs.shape[1]).int()
    # pre-computation, call for call (decode_utils.py:417-438)
    phn = torch.sigmoid(out)
    log_p_yx = log(torch.stack([phn, 1 - phn], dim=3), device_log)
    prior = prior.to(dev) if device_log else prior
    log_p_y = log(torch.stack([prior, 1 - prior], dim=1), device_log)
    boundary_v = predictions["boundary_v"]
    log_p_b = log(torch.stack([boundary_v, 1 - boundary_v], dim=2), device_log)
    log_p_pi = log(torch.softmax(predictions["pi_logits"], dim=-1), device_log)

    boundary, frames, phones, status = decode_from_logs(log_p_yx, log_p_b, log_p_pi, log_p_y, plvl_cnnl_seqs, t_abs, l_abs,
                                                        weight=weight, device=dev)
    boundary, frames, phones, status = boundary.cpu().numpy(), frames.cpu().numpy(), phones.cpu().numpy(), status.cpu().numpy()
    t_abs, l_abs = t_abs.cpu().numpy(), l_abs.cpu().numpy()
    decoded_boundary_seqs, flvl_md_lbl_seqs, plvl_md_lbl_seqs = [], [], []
    for i in range(len(utt_ids)):
        if status[i] != 0:
            # the reference dies in its final `assert l == t == 0` (or on an index error) for these inputs
            raise AssertionError(f"utterance {utt_ids[i]}: no alignment of {int(l_abs[i])} phonemes to {int(t_abs[i])} frames "
                                 f"(status {int(status[i])})")
        decoded_boundary_seqs.append(boundary[i, :t_abs[i]].astype(np.int64))
        flvl_md_lbl_seqs.append([int(v) for v in frames[i, :t_abs[i]]])
        plvl_md_lbl_seqs.append([int(v) for v in phones[i, :l_abs[i]]])
    return decoded_boundary_seqs, flvl_md_lbl_seqs, plvl_md_lbl_seqs
