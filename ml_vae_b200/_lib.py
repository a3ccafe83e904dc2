"""ctypes binding of libmlvae_b200.so (include/mlvae_b200.h).

The library is the product; there is no Python/torch fallback.  Loading fails
loudly when the .so is missing, and every compute entry point raises when it is
handed a non-CUDA tensor.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmlvae_b200.so")

F32, BF16 = 0, 1
RED = {"mean": 0, "batchmean": 1, "batch": 2}
RECON = {"likelihood": 0, "mse": 1}

_vp, _i, _i64, _u64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_size_t

class GemmArgs(C.Structure):
    """mlvae_gemm_args (include/mlvae_b200.h)."""
    _fields_ = [("nprob", _i), ("A", _vp * 4), ("B", _vp * 4), ("D", _vp * 4), ("bias", _vp * 4),
                ("M", _i), ("N", _i), ("K", _i), ("kbatches", _i), ("a_mn_major", _i), ("b_mn_major", _i),
                ("lda", _i64), ("ldb", _i64), ("a_batch_stride", _i64), ("b_batch_stride", _i64), ("ldd", _i64),
                ("out_f32", _i), ("accumulate", _i), ("leaky", _i), ("row_perm_H", _i), ("split_k", _i), ("ws", _vp),
                ("drop_p", C.c_float), ("drop_seed", _u64), ("drop_offset", _u64), ("drop_offset_add", _vp), ("bn", _i), ("max_ctas", _i)]


class ChainFwdArgs(C.Structure):
    """mlvae_chain_fwd_args (include/mlvae_b200.h)."""
    _fields_ = [("nprob", _i), ("x", _vp * 2), ("w_a", _vp * 2), ("w_b", _vp * 2), ("bias_a", _vp * 2), ("bias_b", _vp * 2),
                ("y_a", _vp * 2), ("y_b", _vp * 2), ("M", _i), ("K_A", _i), ("N_A", _i), ("N_B", _i), ("act_b", _i),
                ("ld_x", _i64), ("ld_ya", _i64), ("ld_yb", _i64)]


class ChainBwdArgs(C.Structure):
    """mlvae_chain_bwd_args (include/mlvae_b200.h)."""
    _fields_ = [("nprob", _i), ("g_out", _vp * 2), ("y_b", _vp * 2), ("y_a", _vp * 2), ("x", _vp * 2), ("w_a", _vp * 2), ("w_b", _vp * 2),
                ("dw_a", _vp * 2), ("db_a", _vp * 2), ("dw_b", _vp * 2), ("db_b", _vp * 2), ("dx", _vp * 2),
                ("M", _i), ("K_A", _i), ("N_A", _i), ("N_B", _i), ("act_b", _i),
                ("ld_g", _i64), ("ld_yb", _i64), ("ld_ya", _i64), ("ld_x", _i64), ("ld_dx", _i64), ("ws", _vp)]


class DpAdamArgs(C.Structure):
    """mlvae_dp_adam_args (include/mlvae_b200.h)."""
    _fields_ = [("world", _i), ("rank", _i), ("grads", _vp * 8), ("params", _vp * 8), ("params_bf16", _vp * 8), ("sync", _vp * 8),
                ("mc_grads", _vp), ("mc_params", _vp), ("mc_params_bf16", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp), ("n", _i64),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("max_grad_norm", C.c_float), ("loss", _vp)]


# name -> (restype, argtypes); mirrors include/mlvae_b200.h one to one
SIGNATURES = {
    "mlvae_abi_version": (_i, []),
    "mlvae_last_error": (C.c_char_p, []),
    "mlvae_device_info": (_i, [C.POINTER(_i), C.POINTER(_i)]),
    "mlvae_reduce_scratch_bytes": (_sz, []),
    "mlvae_philox_u32": (_i, [_u64, _u64, _i64, _vp, _vp]),
    "mlvae_philox_normal": (_i, [_u64, _u64, _i64, _vp, _i, _vp]),
    "mlvae_philox_normal_ex": (_i, [_u64, _u64, _i64, _vp, _i, _i, _vp]),
    "mlvae_reparam_kl_fwd": (_i, [_vp, _vp, _vp, _u64, _u64, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "mlvae_reparam_kl_bwd": (_i, [_vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "mlvae_reparam_kl_fwd_strided": (_i, [_vp, _vp, _i64, _vp, _u64, _u64, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "mlvae_reparam_kl_bwd_strided": (_i, [_vp, _vp, _i64, _vp, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i64, _vp]),
    "mlvae_gmm_reparam_kl_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _u64, _u64, _vp, _i64, _i, _vp, _vp, _vp]),
    "mlvae_gmm_reparam_kl_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "mlvae_apply_weight_fwd": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp, _vp]),
    "mlvae_apply_weight_bwd": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _vp]),
    "mlvae_recon_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "mlvae_recon_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "mlvae_masked_reduce_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "mlvae_masked_reduce_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "mlvae_fbank_plan_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _vp, _vp]),
    "mlvae_fbank_plan_destroy": (_i, [_vp]),
    "mlvae_fbank_frames": (_i, [_vp, _i64, _i]),
    "mlvae_fbank_feature_dim": (_i, [_vp]),
    "mlvae_fbank_scratch_bytes": (_sz, [_vp, _i, _i64]),
    "mlvae_debug_set_profile_buffer": (_i, [_vp]),
    "mlvae_debug_set_option": (_i, [_i, _i]),
    "mlvae_lstm_scratch_bytes": (_sz, [_i, _i]),
    "mlvae_lstm_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "mlvae_lstm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "mlvae_lstm_scratch_bytes_dirs": (_sz, [_i, _i, _i]),
    "mlvae_lstm_fwd_dirs": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "mlvae_lstm_bwd_dirs": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "mlvae_norm_state_bytes": (_sz, [_i]),
    "mlvae_norm_scratch_bytes": (_sz, [_i, _i]),
    "mlvae_global_norm": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "mlvae_global_norm_batch_avg": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "mlvae_global_norm_from_avg": (_i, [_vp, _i, _i, _i, _vp, C.c_float, _i, _vp, _vp, _i, _vp]),
    "mlvae_linear_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mlvae_dense_bwd_scratch_bytes": (C.c_size_t, [_i]),
    "mlvae_dense_bwd_prep": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i64, C.c_float, _vp, _i, _vp]),
    "mlvae_adam_state_bytes": (_sz, []),
    "mlvae_dp_sync_bytes": (_sz, []),
    "mlvae_dp_adam_step": (_i, [C.POINTER(DpAdamArgs), _vp]),
    "mlvae_dp_read_state": (_i, [_vp, C.POINTER(C.c_float * 5), _vp]),
    "mlvae_dp_set_adam_step": (_i, [_vp, C.c_float, _vp]),
    "mlvae_dp_debug_max_ctas": (_i, [_i]),
    "mlvae_adam_clip_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, C.c_float, C.c_double, C.c_double, C.c_double, C.c_double, C.c_float, _vp, _vp, _vp]),
    "mlvae_lstm_pack_weights": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "mlvae_lstm_bias_grads": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "mlvae_lstm_unpack_grads": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "mlvae_pcm_unpack": (_i, [_vp, _i, _vp, _vp, _i, _i64, C.c_float, _vp, _vp]),
    "mlvae_tc05_selftest": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "mlvae_gemm_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mlvae_gemm_bf16": (_i, [C.POINTER(GemmArgs), _vp]),
    "mlvae_mlp_chain_fwd": (_i, [C.POINTER(ChainFwdArgs), _vp]),
    "mlvae_mlp_chain_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "mlvae_mlp_chain_bwd": (_i, [C.POINTER(ChainBwdArgs), _vp]),
    "mlvae_dropout": (_i, [_vp, _vp, _i64, C.c_float, _u64, _u64, _vp, _i, _vp]),
    "mlvae_md_decode_workspace_bytes": (_sz, [_i, _i, _i]),
    "mlvae_md_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, C.c_double, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mlvae_fbank_fwd": (_i, [_vp, _vp, _vp, _i, _i64, _i64, _i, _vp, _i, _i, _vp, _vp, _vp]),
}

_lib = None
LAUNCHES = 0      # kernels of libmlvae_b200 launched so far (bench.py reads and resets it)


class MlvaeError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and type the shared library.  No fallback: a missing library is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MlvaeError(
                f"{LIB_PATH} is missing: build it with `python -m ml_vae_b200.build` "
                "(nvcc, sm_100a).  ml_vae_b200 has no CPU or PyTorch fallback.")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name, None)
            if fn is None:
                continue        # optional extension headers (gemm / lstm) bind their own symbols
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def check(rc: int, what: str = "", kernels: int = 1):
    global LAUNCHES
    LAUNCHES += kernels
    if rc != 0:
        msg = lib().mlvae_last_error()
        raise MlvaeError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise MlvaeError(f"unsupported dtype {t.dtype}: the B200 kernels take float32 or bfloat16")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MlvaeError("ml_vae_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on "
                             f"{t.device}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_scratch = {}


def reduce_scratch(device, slot: int = 0) -> torch.Tensor:
    """Zero-initialised, self-resetting reduction scratch; one per (device, stream, slot)."""
    key = (device, torch.cuda.current_stream(device).cuda_stream, slot)
    buf = _scratch.get(key)
    if buf is None:
        buf = torch.zeros(lib().mlvae_reduce_scratch_bytes(), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf
