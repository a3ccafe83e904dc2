"""Dense (Linear / LeakyReLU) stacks of the latent block.

Reference: modules/fc_block.py:4-21 (Linear -> LeakyReLU(0.01) ... last Linear bare,
optional end activation; the ``dropout`` argument is accepted and ignored there).

Round-1 state: the projections are library GEMMs (cuBLAS through torch.nn.functional.linear)
-- they are HBM-bound skinny GEMMs (N <= 128) -- while the tcgen05/TMEM fused chain
(csrc/gemm_chain.cu) is brought up; `linear_chain` is the single seam both go through.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib as L

LEAKY_SLOPE = 0.01


def linear_chain(x: torch.Tensor, weights, biases, end_activation: bool = False) -> torch.Tensor:
    """x (..., K0) -> (..., N_last).  weights[i]: (N_i, K_i) float32 master copies."""
    L.require_cuda(x)
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = F.linear(x, w.to(x.dtype), b.to(x.dtype))
        if i + 1 < n or end_activation:
            x = F.leaky_relu(x, LEAKY_SLOPE)
    return x
