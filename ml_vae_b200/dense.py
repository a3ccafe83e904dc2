"""Dense (Linear / LeakyReLU) stacks of the latent block.

Reference: modules/fc_block.py:4-21 (Linear -> LeakyReLU(0.01) ... last Linear bare, optional end
activation; the ``dropout`` argument is accepted and ignored there).

bf16 activations run every Linear(+LeakyReLU) on the tcgen05 / TMEM kernel ``mlvae_linear_fwd``
(csrc/gemm_chain.cu): forward and the input gradient (the same kernel against W^T); the weight
gradient dW = g^T x, a reduction over all B*T rows, runs on the TMA / tcgen05 GEMM (csrc/gemm.cu, split-K).
The float32 stack stays on library GEMMs (bf16 tensor cores would break the fp32 1e-5 parity).  ``linear_chain`` is the single
seam every module goes through.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib as L

LEAKY_SLOPE = 0.01


def _tc_ok(x2: torch.Tensor, n_out: int, k_in: int) -> bool:
    return (x2.dtype == torch.bfloat16 and n_out <= 256 and k_in % 8 == 0 and x2.stride(0) % 8 == 0
            and x2.data_ptr() % 16 == 0)


def _launch(x2, w_bf16, bias_f32, n_out, leaky):
    M, K = x2.shape
    y = torch.empty(M, n_out, dtype=torch.bfloat16, device=x2.device)
    L.check(L.lib().mlvae_linear_fwd(L.ptr(x2), L.ptr(w_bf16), L.ptr(bias_f32), L.ptr(y), M, n_out, K, x2.stride(0), n_out,
                                     int(leaky), L.stream_ptr()), "mlvae_linear_fwd")
    return y


_db_scratch = {}


def _bwd_prep(dy: torch.Tensor, y, accumulate_into=None):
    """g = dy * leaky'(y) (y given) and db = column sums of g, one kernel (csrc/dense_bwd.cu).  ``accumulate_into``: a float32
    (N,) tensor that receives db += ... instead of a fresh one."""
    M, N = dy.shape
    key = (dy.device, torch.cuda.current_stream(dy.device).cuda_stream)
    need = L.lib().mlvae_dense_bwd_scratch_bytes(N)
    buf = _db_scratch.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.zeros(max(need, L.lib().mlvae_dense_bwd_scratch_bytes(256)), dtype=torch.uint8, device=dy.device)
        _db_scratch[key] = buf
    g = torch.empty_like(dy) if y is not None else None
    db = accumulate_into if accumulate_into is not None else torch.empty(N, dtype=torch.float32, device=dy.device)
    L.check(L.lib().mlvae_dense_bwd_prep(L.ptr(dy), L.ptr(y), L.ptr(g), L.ptr(db), M, N, N, LEAKY_SLOPE, L.ptr(buf), int(accumulate_into is not None),
                                         L.stream_ptr()), "mlvae_dense_bwd_prep")
    return (g if y is not None else dy), db


def _weight_grad(g: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """dW = g^T x (float32) on the TMA / tcgen05 GEMM (csrc/gemm.cu): both operands MN-major (no transposes), the reduction over
    the B*T rows split across the SMs with a deterministic second pass.  Odd shapes fall back to a library GEMM."""
    M, N = g.shape
    K = x2.shape[1]
    if (N % 8 == 0 and K % 8 == 0 and g.stride(1) == 1 and x2.stride(1) == 1 and g.stride(0) % 8 == 0 and x2.stride(0) % 8 == 0
            and g.data_ptr() % 16 == 0 and x2.data_ptr() % 16 == 0 and g.dtype == torch.bfloat16 and x2.dtype == torch.bfloat16):
        from .gemm import gemm
        dw = torch.empty(N, K, dtype=torch.float32, device=g.device)
        tiles = ((N + 127) // 128) * ((K + 255) // 256)
        split = max(1, min(32, 128 // tiles, (M + 1023) // 1024))
        gemm(g, x2, dw, N, K, M, lda=g.stride(0), ldb=x2.stride(0), ldd=K, a_mn=True, b_mn=True, out_f32=True, split_k=split)
        return dw
    return torch.mm(g.t(), x2, out_dtype=torch.float32)


class _LinearTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, w, b, leaky):
        wb = w.detach().to(torch.bfloat16).contiguous()
        y = _launch(x2, wb, b.detach().float().contiguous(), w.shape[0], leaky)
        ctx.save_for_backward(x2, wb, y if leaky else None)
        ctx.leaky = leaky
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, wb, y = ctx.saved_tensors
        g = dy.contiguous()
        N, K = wb.shape
        db = None
        if N % 8 == 0 and N <= 2048 and g.dtype == torch.bfloat16:
            g, db = _bwd_prep(g, y if ctx.leaky else None)
        elif ctx.leaky:
            g = g * torch.where(y > 0, 1.0, LEAKY_SLOPE).to(g.dtype)
        dx = None
        if ctx.needs_input_grad[0]:
            if _tc_ok(g, K, N):
                dx = _launch(g, wb.t().contiguous(), None, K, False)      # dx = g W  ==  linear(g, W^T)
            else:
                dx = g @ wb
        dw = _weight_grad(g, x2)
        if db is None:
            db = g.sum(0, dtype=torch.float32)
        return dx, dw, db, None


class _LinearDirect(torch.autograd.Function):
    """Linear(+LeakyReLU) over ARENA views (train_step.FlatArena.linear_views): the bf16 shadow weight and float32 bias are read
    as they are (no cast / cat / transpose kernels), and the backward pass ACCUMULATES dW and db straight into the gradient
    bucket (GEMM epilogue / bias kernel), so autograd sees one differentiable input and adds nothing."""

    @staticmethod
    def forward(ctx, x2, w16, b32, gw, gb, leaky, anchor):
        # `anchor` is the float32 master parameter: only there so that autograd records this node even when x itself does not
        # require a gradient (first layer of the encoder); its gradient is accumulated in place, not returned
        N, K = w16.shape
        if N <= 256:
            y = _launch(x2, w16, b32, N, leaky)                  # skinny output: the one-pass streaming kernel (csrc/gemm_chain.cu)
        else:
            from .gemm import gemm                              # wide output: the tiled TMA GEMM with the same epilogue
            y = torch.empty(x2.shape[0], N, dtype=torch.bfloat16, device=x2.device)
            gemm(x2, w16, y, x2.shape[0], N, K, lda=x2.stride(0), ldb=K, ldd=N, bias=b32, leaky=leaky)
        ctx.save_for_backward(x2, y if leaky else None)
        ctx.w16, ctx.gw, ctx.gb, ctx.leaky = w16, gw, gb, leaky
        return y

    @staticmethod
    def backward(ctx, dy):
        from .gemm import gemm
        x2, y = ctx.saved_tensors
        w16, gw, gb = ctx.w16, ctx.gw, ctx.gb
        N, K = w16.shape
        M = x2.shape[0]
        g = dy if dy.is_contiguous() else dy.contiguous()
        g, _ = _bwd_prep(g, y if ctx.leaky else None, accumulate_into=gb)          # g = dy * act'(y), db accumulated in place
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, dtype=torch.bfloat16, device=g.device)
            gemm(g, w16, dx, M, K, N, lda=N, ldb=K, ldd=K, b_mn=True)                # dx = g W  (W row-major (N, K): MN-major B)
        tiles = ((N + 127) // 128) * ((K + 255) // 256)
        split = max(1, min(32, 128 // tiles, (M + 1023) // 1024))
        gemm(g, x2, gw, N, K, M, lda=N, ldb=x2.stride(0), ldd=K, a_mn=True, b_mn=True, out_f32=True, accumulate=True, split_k=split)
        return dx, None, None, None, None, None, None


def linear_direct(x: torch.Tensor, views, leaky: bool = False) -> torch.Tensor:
    """x (..., K) bf16 -> act(x W^T + b) with ``views`` = (w_bf16, bias_f32, grad_w, grad_b, master parameter) from
    FlatArena.linear_views."""
    L.require_cuda(x)
    w16, b32, gw, gb, anchor = views
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(1) != 1 or x2.stride(0) % 8 or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    y = _LinearDirect.apply(x2, w16, b32, gw, gb, leaky, anchor)
    return y.reshape(*lead, w16.shape[0])


def direct_chain(x: torch.Tensor, views_list, end_activation: bool = False) -> torch.Tensor:
    n = len(views_list)
    for i, v in enumerate(views_list):
        x = linear_direct(x, v, leaky=(i + 1 < n or end_activation))
    return x


def linear(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, leaky: bool = False) -> torch.Tensor:
    """x (..., K) -> act(x W^T + b) (..., N); w, b are the float32 master parameters."""
    L.require_cuda(x)
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if _tc_ok(x2, w.shape[0], w.shape[1]):
        y = _LinearTC.apply(x2 if x2.stride(1) == 1 else x2.contiguous(), w, b, leaky)
    else:
        y = F.linear(x2, w.to(x.dtype), b.to(x.dtype))
        if leaky:
            y = F.leaky_relu(y, LEAKY_SLOPE)
    return y.reshape(*lead, w.shape[0])


def linear_chain(x: torch.Tensor, weights, biases, end_activation: bool = False) -> torch.Tensor:
    """x (..., K0) -> (..., N_last).  weights[i]: (N_i, K_i) float32 master copies."""
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = linear(x, w, b, leaky=(i + 1 < n or end_activation))
    return x
