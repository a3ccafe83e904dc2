"""Fused two-layer Linear / LeakyReLU chains (csrc/mlp_chain.cu, mlvae_mlp_chain_fwd / _bwd): the encoder trunk
(modules/vanilla_vae.py:13-24) and the tails of the decoder heads (modules/decoder.py:16-17,24-25; both heads in one launch).

``chain2`` is the autograd entry point over ARENA views (train_step.FlatArena.linear_views): bf16 shadow weights are read as
they are, weight / bias gradients are accumulated in place inside the kernels, autograd sees only the activations."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_ws = {}


def supported(k_a: int, n_a: int, n_b: int) -> bool:
    return all(v % 16 == 0 for v in (k_a, n_a, n_b)) and k_a <= 112 and n_a <= 112 and n_b <= 128


def _workspace(nprob, k_a, n_a, device):
    need = L.lib().mlvae_mlp_chain_bwd_workspace_bytes(nprob, k_a, n_a)
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def chain_fwd(xs, views_a, views_b, act_b: bool, save_hidden: bool = True):
    """xs: list of 1 or 2 (M, K_A) bf16 matrices (unit column stride, any row stride % 8 == 0).  Returns (y_a list, y_b list)."""
    M, KA = xs[0].shape
    NA, NB = views_a[0][0].shape[0], views_b[0][0].shape[0]
    dev = xs[0].device
    ya = [torch.empty(M, NA, dtype=torch.bfloat16, device=dev) if save_hidden else None for _ in xs]
    yb = [torch.empty(M, NB, dtype=torch.bfloat16, device=dev) for _ in xs]
    a = L.ChainFwdArgs()
    a.nprob = len(xs)
    for i, x in enumerate(xs):
        L.require_cuda(x)
        a.x[i], a.w_a[i], a.w_b[i] = x.data_ptr(), views_a[i][0].data_ptr(), views_b[i][0].data_ptr()
        a.bias_a[i], a.bias_b[i] = views_a[i][1].data_ptr(), views_b[i][1].data_ptr()
        a.y_a[i] = None if ya[i] is None else ya[i].data_ptr()
        a.y_b[i] = yb[i].data_ptr()
    a.M, a.K_A, a.N_A, a.N_B, a.act_b = M, KA, NA, NB, int(act_b)
    a.ld_x, a.ld_ya, a.ld_yb = xs[0].stride(0), NA, NB
    L.check(L.lib().mlvae_mlp_chain_fwd(C.byref(a), L.stream_ptr()), "mlvae_mlp_chain_fwd")
    return ya, yb


def chain_bwd(gs, ybs, yas, xs, views_a, views_b, act_b: bool, dx):
    """Accumulates dW / db of both layers into the gradient views and writes dx (a list of (M, K_A) views, or None)."""
    M, KA = xs[0].shape
    NA, NB = views_a[0][0].shape[0], views_b[0][0].shape[0]
    a = L.ChainBwdArgs()
    a.nprob = len(xs)
    for i in range(len(xs)):
        a.g_out[i], a.y_a[i], a.x[i] = gs[i].data_ptr(), yas[i].data_ptr(), xs[i].data_ptr()
        a.y_b[i] = ybs[i].data_ptr() if act_b else None
        a.w_a[i], a.w_b[i] = views_a[i][0].data_ptr(), views_b[i][0].data_ptr()
        a.dw_a[i], a.db_a[i] = views_a[i][2].data_ptr(), views_a[i][3].data_ptr()
        a.dw_b[i], a.db_b[i] = views_b[i][2].data_ptr(), views_b[i][3].data_ptr()
        a.dx[i] = None if dx is None else dx[i].data_ptr()
    a.M, a.K_A, a.N_A, a.N_B, a.act_b = M, KA, NA, NB, int(act_b)
    a.ld_g, a.ld_yb, a.ld_ya, a.ld_x = gs[0].stride(0), NB, NA, xs[0].stride(0)
    a.ld_dx = 0 if dx is None else dx[0].stride(0)
    ws = _workspace(len(xs), KA, NA, xs[0].device)
    a.ws = ws.data_ptr()
    L.check(L.lib().mlvae_mlp_chain_bwd(C.byref(a), L.stream_ptr()), "mlvae_mlp_chain_bwd", kernels=2)


class _Chain2(torch.autograd.Function):
    """x (M, K_tot) -> y_b.  nprob = 1: one chain over all of x.  nprob = 2: x's columns are split in two halves that go through
    two independent chains (the two decoder heads); the outputs are returned as two tensors."""

    @staticmethod
    def forward(ctx, x2, views_a, views_b, act_b, anchor):
        n = len(views_a)
        KA = x2.shape[1] // n
        xs = [x2[:, i * KA:(i + 1) * KA] for i in range(n)]
        ya, yb = chain_fwd(xs, views_a, views_b, act_b)
        ctx.save_for_backward(x2, *ya, *(yb if act_b else []))
        ctx.views_a, ctx.views_b, ctx.act_b, ctx.n = views_a, views_b, act_b, n
        return tuple(yb)

    @staticmethod
    def backward(ctx, *gys):
        n = ctx.n
        saved = ctx.saved_tensors
        x2, ya = saved[0], list(saved[1:1 + n])
        yb = list(saved[1 + n:]) if ctx.act_b else [None] * n
        KA = x2.shape[1] // n
        xs = [x2[:, i * KA:(i + 1) * KA] for i in range(n)]
        gs = [g if (g.stride(1) == 1 and g.stride(0) % 8 == 0 and g.data_ptr() % 16 == 0) else g.contiguous() for g in gys]
        if len({g.stride(0) for g in gs}) > 1:
            gs = [g.contiguous() for g in gs]
        dx = None
        dxs = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x2)                               # both heads write their half of one (M, K_tot) buffer
            dxs = [dx[:, i * KA:(i + 1) * KA] for i in range(n)]
        chain_bwd(gs, yb, ya, xs, ctx.views_a, ctx.views_b, ctx.act_b, dxs)
        return dx, None, None, None, None


def chain2(x: torch.Tensor, views_a, views_b, act_b: bool):
    """x (..., n * K_A) bf16 through n = len(views_a) two-layer chains; returns a tuple of n outputs (..., N_B)."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(1) != 1 or x2.stride(0) % 8 or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    outs = _Chain2.apply(x2, views_a, views_b, act_b, views_a[0][4])
    return tuple(o.reshape(*lead, o.shape[-1]) for o in outs)
