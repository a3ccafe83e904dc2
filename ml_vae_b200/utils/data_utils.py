"""Drop-in for the hot-path part of the reference's utils/data_utils.py.

apply_lens_to_loss (data_utils.py:67-104): same signature and reductions, one CUDA
kernel forward and one backward instead of ones_like / mask build / two multiplies /
two reductions.  The frame predicate is the reference's float32 expression
``arange(T) < lens * T`` evaluated literally (no rounding).
"""
from __future__ import annotations

from .. import ops


def apply_lens_to_loss(loss, lens, reduction="mean"):
    if reduction not in ("mean", "batchmean", "batch"):
        # the reference silently returns the masked tensor here; refusing is safer than
        # pretending to support an unreduced masked output nobody calls
        raise ValueError(f"Invalid reduction: {reduction}")
    return ops.masked_reduce(loss, lens, reduction)


def apply_weight(x, weight):
    """utils/data_utils.py:32-64: x (B,T,N,C) or (B,T,N*C), weight (B,T,N) -> (B,T,C) on one CUDA kernel."""
    return ops.apply_weight(x, weight)
