"""Host wrapper of the TMA-fed tcgen05 GEMM (csrc/gemm.cu, include/mlvae_b200.h: mlvae_gemm_bf16).

Replaces the cuBLAS calls torch made for the time-parallel products of the step: the LSTM input projection and its
input / weight gradients (modules/decoder.py:14-15,22) and the weight gradients / wide forward layers of the FC stacks
(modules/fc_block.py:9-16).  Tensors are passed as (possibly offset) views; only their data pointers and the explicit
leading dimensions are used, so row-shifted and column-sliced operands need no copies.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_ws = {}
PROBE = None      # bench.py sets this to a list; every GEMM launch then appends (flops, start_event, end_event)


def _workspace(nbytes: int, device):
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 22), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def gemm(A, B, D, M: int, N: int, K: int, *, lda: int, ldb: int, ldd: int, a_mn: bool = False, b_mn: bool = False,
         kbatches: int = 1, a_batch_stride: int = 0, b_batch_stride: int = 0, bias=None, out_f32: bool = False,
         accumulate: bool = False, leaky: bool = False, row_perm_H: int = 0, split_k: int = 1, drop_p: float = 0.0,
         drop_seed: int = 0, drop_offset: int = 0, drop_offset_dev=None, bn: int = 0, max_ctas: int = 0):
    """D_i (+)= epilogue(A_i B_i) for the problems i of the lists A, B, D (single tensors are wrapped).  See
    include/mlvae_b200.h for the operand conventions (K-major / MN-major, batched reduction, epilogue)."""
    As, Bs, Ds = (list(t) if isinstance(t, (list, tuple)) else [t] for t in (A, B, D))
    biases = list(bias) if isinstance(bias, (list, tuple)) else [bias] * len(As)
    a = L.GemmArgs()
    a.nprob = len(As)
    for i, (x, y, z, b) in enumerate(zip(As, Bs, Ds, biases)):
        L.require_cuda(x, y, z)
        a.A[i], a.B[i], a.D[i] = x.data_ptr(), y.data_ptr(), z.data_ptr()
        a.bias[i] = None if b is None else b.data_ptr()
    a.M, a.N, a.K, a.kbatches = M, N, K, kbatches
    a.a_mn_major, a.b_mn_major = int(a_mn), int(b_mn)
    a.lda, a.ldb, a.ldd = lda, ldb, ldd
    a.a_batch_stride, a.b_batch_stride = a_batch_stride, b_batch_stride
    a.out_f32, a.accumulate, a.leaky, a.row_perm_H = int(out_f32), int(accumulate), int(leaky), row_perm_H
    a.split_k = split_k
    ws = None
    if split_k > 1:
        ws = _workspace(L.lib().mlvae_gemm_workspace_bytes(len(As), M, N, split_k), Ds[0].device)
        a.ws = ws.data_ptr()
    a.drop_p, a.drop_seed, a.drop_offset = float(drop_p), drop_seed, drop_offset
    a.drop_offset_add = None if drop_offset_dev is None else drop_offset_dev.data_ptr()
    a.bn = bn
    a.max_ctas = max_ctas
    ev0 = None
    if PROBE is not None:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev0.record()
    L.check(L.lib().mlvae_gemm_bf16(C.byref(a), L.stream_ptr()), "mlvae_gemm_bf16", kernels=2 if split_k > 1 else 1)
    if ev0 is not None:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        PROBE.append((2.0 * len(As) * M * N * K * kbatches, ev0, ev1))
    return D
