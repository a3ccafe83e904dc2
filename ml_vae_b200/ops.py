"""torch.autograd.Function wrappers over the C-ABI kernels (include/mlvae_b200.h).

torch is used for device memory, streams and autograd plumbing only; all
arithmetic of these ops happens in libmlvae_b200.so.  Reference semantics:
  reparam_kl      modules/vanilla_vae.py:37-45  (+ utils/data_utils.py:67-104 when reduced)
  recon_loss      modules/decoder.py:37-53      (+ utils/data_utils.py:67-104 when reduced)
  masked_reduce   utils/data_utils.py:67-104
  philox_normal   replaces torch.randn_like at modules/vanilla_vae.py:39
"""
from __future__ import annotations

import torch

from . import _lib as L


def _c(t):
    return None if t is None else t.contiguous()


def _btc(t: torch.Tensor):
    if t.dim() < 2:
        raise ValueError(f"expected (B, T, ...) tensor, got shape {tuple(t.shape)}")
    b, tt = t.shape[0], t.shape[1]
    c = 1
    for s in t.shape[2:]:
        c *= s
    return b, tt, c


def _lens(lens: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return lens.to(device=like.device, dtype=torch.float32).contiguous()


def philox_normal(shape, seed: int, offset: int = 0, dtype=torch.float32, device="cuda", kernel_dtype=None) -> torch.Tensor:
    """Materialise the eps stream the fused kernels draw for (seed, offset).  The float32 and the bf16 kernels draw different
    streams (4 vs 8 normals per Philox call, csrc/philox.cuh): ``kernel_dtype`` (default: ``dtype``) says whose; e.g.
    ``philox_normal(shape, s, o, kernel_dtype=torch.bfloat16)`` = float32 values of what the bf16 kernels used."""
    out = torch.empty(shape, dtype=dtype, device=device)
    L.require_cuda(out)
    kd = L.dtype_code(out) if kernel_dtype is None else (L.BF16 if kernel_dtype == torch.bfloat16 else L.F32)
    L.check(L.lib().mlvae_philox_normal_ex(seed, offset, out.numel(), L.ptr(out), L.dtype_code(out), kd, L.stream_ptr()),
            "mlvae_philox_normal")
    return out


def philox_u32(n: int, seed: int, offset: int = 0, device="cuda") -> torch.Tensor:
    out = torch.empty(n, dtype=torch.int32, device=device)
    L.check(L.lib().mlvae_philox_u32(seed, offset, n, L.ptr(out), L.stream_ptr()), "mlvae_philox_u32")
    return out


class _ReparamKL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, eps, lens, seed, offset, want_elem, want_mean, offset_dev=None):
        L.require_cuda(mu, logvar, eps)
        if mu.shape != logvar.shape or mu.dtype != logvar.dtype:
            raise ValueError("mean and log_var must have the same shape and dtype")
        mu, logvar, eps = _c(mu), _c(logvar), _c(eps)
        if eps is not None and (eps.shape != mu.shape or eps.dtype != mu.dtype):
            eps = eps.to(mu.dtype).reshape(mu.shape).contiguous()
        B, T, C = _btc(mu)
        z = torch.empty_like(mu)
        kl_elem = torch.empty_like(mu) if want_elem else None
        kl_out = torch.empty(3, dtype=torch.float32, device=mu.device) if want_mean else None
        lens_f = _lens(lens, mu) if lens is not None else None
        if want_mean and lens_f is None:
            raise ValueError("reduced KL needs lens")
        L.check(L.lib().mlvae_reparam_kl_fwd(
            L.ptr(mu), L.ptr(logvar), L.ptr(eps), seed, offset, L.ptr(offset_dev), L.ptr(lens_f), B, T, C, L.dtype_code(mu),
            L.ptr(z), L.ptr(kl_elem), L.ptr(kl_out), L.ptr(L.reduce_scratch(mu.device)) if want_mean else None,
            L.stream_ptr()), "mlvae_reparam_kl_fwd")
        ctx.save_for_backward(mu, logvar, eps, lens_f)
        ctx.seed, ctx.offset, ctx.offset_dev = seed, offset, offset_dev
        ctx.set_materialize_grads(False)
        empty = mu.new_empty(0)
        return z, (kl_elem if want_elem else empty), (kl_out[0] if want_mean else empty.float())

    @staticmethod
    def backward(ctx, gz, gelem, gmean):
        mu, logvar, eps, lens_f = ctx.saved_tensors
        B, T, C = _btc(mu)
        gz = _c(gz)
        gelem = _c(gelem) if gelem is not None and gelem.numel() else None
        gmean = gmean.contiguous().float() if gmean is not None and gmean.numel() else None
        if gz is not None and gz.dtype != mu.dtype:
            gz = gz.to(mu.dtype)
        gmu, glv = torch.empty_like(mu), torch.empty_like(mu)
        L.check(L.lib().mlvae_reparam_kl_bwd(
            L.ptr(mu), L.ptr(logvar), L.ptr(eps), ctx.seed, ctx.offset, L.ptr(ctx.offset_dev), L.ptr(gz), L.ptr(gelem), L.ptr(gmean),
            L.ptr(lens_f), B, T, C, L.dtype_code(mu), L.ptr(gmu), L.ptr(glv), L.stream_ptr()), "mlvae_reparam_kl_bwd")
        return gmu, glv, None, None, None, None, None, None, None


class _ReparamKLStacked(torch.autograd.Function):
    """reparam + KL reading mean | log_var as the two halves of ONE (B, T, 2L) projection output (the stacked-head GEMM of
    modules/vanilla_vae.py:23-24) in place, and writing both gradients into one (B, T, 2L) buffer that feeds that GEMM's
    backward: no slice copies forward, no zero-fill / scatter / add of slice gradients backward."""

    @staticmethod
    def forward(ctx, ml, eps, lens, seed, offset, want_elem, want_mean, offset_dev=None):
        L.require_cuda(ml, eps)
        ml = _c(ml)
        B, T, C2 = _btc(ml)
        if C2 % 2:
            raise ValueError("stacked mean | log_var needs an even channel count")
        C = C2 // 2
        shape = tuple(ml.shape[:-1]) + (C,)
        if eps is not None:
            eps = eps.to(ml.dtype).reshape(shape).contiguous()
        z = torch.empty(shape, dtype=ml.dtype, device=ml.device)
        kl_elem = torch.empty_like(z) if want_elem else None
        kl_out = torch.empty(3, dtype=torch.float32, device=ml.device) if want_mean else None
        lens_f = _lens(lens, ml) if lens is not None else None
        if want_mean and lens_f is None:
            raise ValueError("reduced KL needs lens")
        half = C * ml.element_size()
        L.check(L.lib().mlvae_reparam_kl_fwd_strided(
            ml.data_ptr(), ml.data_ptr() + half, C2, L.ptr(eps), seed, offset, L.ptr(offset_dev), L.ptr(lens_f), B, T, C,
            L.dtype_code(ml), L.ptr(z), L.ptr(kl_elem), L.ptr(kl_out), L.ptr(L.reduce_scratch(ml.device)) if want_mean else None,
            L.stream_ptr()), "mlvae_reparam_kl_fwd_strided")
        ctx.save_for_backward(ml, eps, lens_f)
        ctx.seed, ctx.offset, ctx.offset_dev = seed, offset, offset_dev
        ctx.set_materialize_grads(False)
        empty = ml.new_empty(0)
        return z, (kl_elem if want_elem else empty), (kl_out[0] if want_mean else empty.float())

    @staticmethod
    def backward(ctx, gz, gelem, gmean):
        ml, eps, lens_f = ctx.saved_tensors
        B, T, C2 = _btc(ml)
        C = C2 // 2
        gz = _c(gz)
        gelem = _c(gelem) if gelem is not None and gelem.numel() else None
        gmean = gmean.contiguous().float() if gmean is not None and gmean.numel() else None
        if gz is not None and gz.dtype != ml.dtype:
            gz = gz.to(ml.dtype)
        if gelem is not None and gelem.dtype != ml.dtype:
            gelem = gelem.to(ml.dtype)
        gml = torch.empty_like(ml)
        half = C * ml.element_size()
        L.check(L.lib().mlvae_reparam_kl_bwd_strided(
            ml.data_ptr(), ml.data_ptr() + half, C2, L.ptr(eps), ctx.seed, ctx.offset, L.ptr(ctx.offset_dev), L.ptr(gz), L.ptr(gelem),
            L.ptr(gmean), L.ptr(lens_f), B, T, C, L.dtype_code(ml), gml.data_ptr(), gml.data_ptr() + half, C2, L.stream_ptr()),
            "mlvae_reparam_kl_bwd_strided")
        return gml, None, None, None, None, None, None, None


def reparam_kl_stacked(ml, lens=None, eps=None, seed: int = 0, offset: int = 0,
                       want_elem: bool = False, want_mean: bool = True, offset_dev=None):
    """``reparam_kl`` on ml = [mean | log_var] (B, T, 2L), read in place -> (z, kl_elem | None, kl_mean | None)."""
    z, e, m = _ReparamKLStacked.apply(ml, eps, lens, int(seed), int(offset), want_elem, want_mean, offset_dev)
    return z, (e if want_elem else None), (m if want_mean else None)


def reparam_kl(mu, logvar, lens=None, eps=None, seed: int = 0, offset: int = 0,
               want_elem: bool = False, want_mean: bool = True, offset_dev=None):
    """-> (z, kl_elem | None, kl_mean | None).  eps=None draws Philox(seed, offset [+ offset_dev[0], a device
    int64 step counter that keeps CUDA-graph replays drawing fresh noise])."""
    z, e, m = _ReparamKL.apply(mu, logvar, eps, lens, int(seed), int(offset), want_elem, want_mean, offset_dev)
    return z, (e if want_elem else None), (m if want_mean else None)


class _GmmReparamKL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, pmu, plogvar, eps, seed, offset):
        L.require_cuda(mu, logvar, pmu, plogvar, eps)
        mu, logvar, pmu, plogvar, eps = _c(mu), _c(logvar), _c(pmu), _c(plogvar), _c(eps)
        if not (mu.shape == logvar.shape == pmu.shape == plogvar.shape):
            raise ValueError("mean, log_var, prior_mean, prior_log_var must have the same shape")
        if eps is not None:
            eps = eps.to(mu.dtype).reshape(mu.shape).contiguous()
        z, kl = torch.empty_like(mu), torch.empty_like(mu)
        L.check(L.lib().mlvae_gmm_reparam_kl_fwd(L.ptr(mu), L.ptr(logvar), L.ptr(pmu), L.ptr(plogvar), L.ptr(eps), seed, offset,
                                                 None, mu.numel(), L.dtype_code(mu), L.ptr(z), L.ptr(kl), L.stream_ptr()),
                "mlvae_gmm_reparam_kl_fwd")
        ctx.save_for_backward(mu, logvar, pmu, plogvar, eps)
        ctx.seed, ctx.offset = seed, offset
        ctx.set_materialize_grads(False)
        return z, kl

    @staticmethod
    def backward(ctx, gz, gk):
        mu, logvar, pmu, plogvar, eps = ctx.saved_tensors
        gz = _c(gz.to(mu.dtype)) if gz is not None else None
        gk = _c(gk.to(mu.dtype)) if gk is not None else None
        outs = [torch.empty_like(mu) for _ in range(4)]
        L.check(L.lib().mlvae_gmm_reparam_kl_bwd(L.ptr(mu), L.ptr(logvar), L.ptr(pmu), L.ptr(plogvar), L.ptr(eps), ctx.seed,
                                                 ctx.offset, None, L.ptr(gz), L.ptr(gk), mu.numel(), L.dtype_code(mu),
                                                 *[L.ptr(o) for o in outs], L.stream_ptr()), "mlvae_gmm_reparam_kl_bwd")
        return outs[0], outs[1], outs[2], outs[3], None, None, None


def gmm_reparam_kl(mu, logvar, prior_mu, prior_logvar, eps=None, seed: int = 0, offset: int = 0):
    """GMMVAE.reparameterize + compute_kld_loss (modules/gmm_vae.py:51-67) -> (z, unreduced kl)."""
    return _GmmReparamKL.apply(mu, logvar, prior_mu, prior_logvar, eps, int(seed), int(offset))


class _ApplyWeight(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        L.require_cuda(x, w)
        B, T, N = w.shape
        C = x.shape[-1] // N if x.dim() == 3 else x.shape[-1]
        x = x.reshape(B * T, N, C).contiguous()
        w = w.reshape(B * T, N).to(x.dtype).contiguous()
        out = torch.empty(B * T, C, dtype=x.dtype, device=x.device)
        L.check(L.lib().mlvae_apply_weight_fwd(L.ptr(x), L.ptr(w), B * T, N, C, L.dtype_code(x), L.ptr(out), L.stream_ptr()),
                "mlvae_apply_weight_fwd")
        ctx.save_for_backward(x, w)
        ctx.meta = (B, T, N, C)
        return out.view(B, T, C)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        B, T, N, C = ctx.meta
        g = g.reshape(B * T, C).to(x.dtype).contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        L.check(L.lib().mlvae_apply_weight_bwd(L.ptr(x), L.ptr(w), L.ptr(g), B * T, N, C, L.dtype_code(x), L.ptr(gx), L.ptr(gw),
                                               L.stream_ptr()), "mlvae_apply_weight_bwd")
        return (gx.view(ctx.x_shape) if gx is not None else None), (gw.view(B, T, N) if gw is not None else None)


def apply_weight(x, weight):
    """utils/data_utils.py:32-64: x (B,T,N,C) or (B,T,N*C), weight (B,T,N) -> (B,T,C)."""
    return _ApplyWeightShaped.apply(x, weight, tuple(x.shape))


class _ApplyWeightShaped(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, shape):
        ctx.x_shape = shape
        return _ApplyWeight.forward(ctx, x, w)

    @staticmethod
    def backward(ctx, g):
        gx, gw = _ApplyWeight.backward(ctx, g)
        return gx, gw, None


class _ReconLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, logvar, target, lens, loss_type, want_elem, want_mean):
        if loss_type not in L.RECON:
            raise ValueError(f"Invalid loss type: {loss_type}")          # decoder.py:51
        L.require_cuda(mean, logvar, target)
        mean, logvar = _c(mean), _c(logvar)
        target = _c(target.to(mean.dtype))
        if target.shape != mean.shape:
            raise ValueError("target and mean must have the same shape")
        B, T, C = _btc(mean)
        elem = torch.empty_like(mean) if want_elem else None
        out = torch.empty(3, dtype=torch.float32, device=mean.device) if want_mean else None
        lens_f = _lens(lens, mean) if lens is not None else None
        if want_mean and lens_f is None:
            raise ValueError("reduced reconstruction loss needs lens")
        L.check(L.lib().mlvae_recon_fwd(
            L.ptr(mean), L.ptr(logvar), L.ptr(target), L.ptr(lens_f), B, T, C, L.dtype_code(mean), L.RECON[loss_type],
            L.ptr(elem), L.ptr(out), L.ptr(L.reduce_scratch(mean.device)) if want_mean else None, L.stream_ptr()),
            "mlvae_recon_fwd")
        ctx.save_for_backward(mean, logvar, target, lens_f)
        ctx.loss_type = loss_type
        ctx.set_materialize_grads(False)
        empty = mean.new_empty(0)
        return (elem if want_elem else empty), (out[0] if want_mean else empty.float())

    @staticmethod
    def backward(ctx, gelem, gmean):
        mean, logvar, target, lens_f = ctx.saved_tensors
        B, T, C = _btc(mean)
        gelem = _c(gelem) if gelem is not None and gelem.numel() else None
        gmean = gmean.contiguous().float() if gmean is not None and gmean.numel() else None
        gm = torch.empty_like(mean)
        glv = torch.empty_like(mean) if ctx.loss_type == "likelihood" else None
        gt = torch.empty_like(mean) if ctx.needs_input_grad[2] else None
        L.check(L.lib().mlvae_recon_bwd(
            L.ptr(mean), L.ptr(logvar), L.ptr(target), L.ptr(gelem), L.ptr(gmean), L.ptr(lens_f), B, T, C,
            L.dtype_code(mean), L.RECON[ctx.loss_type], L.ptr(gm), L.ptr(glv), L.ptr(gt), L.stream_ptr()),
            "mlvae_recon_bwd")
        return gm, glv, gt, None, None, None, None


def recon_loss(mean, logvar, target, lens=None, loss_type: str = "likelihood",
               want_elem: bool = False, want_mean: bool = True):
    """-> (elem | None, masked mean | None)."""
    e, m = _ReconLoss.apply(mean, logvar, target, lens, loss_type, want_elem, want_mean)
    return (e if want_elem else None), (m if want_mean else None)


class _MaskedReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loss, lens, reduction):
        if reduction not in L.RED:
            raise ValueError(f"Invalid reduction: {reduction}")
        L.require_cuda(loss)
        loss = _c(loss)
        B, T, C = _btc(loss)
        lens_f = _lens(lens, loss)
        out = torch.empty(B if reduction == "batch" else 1, dtype=torch.float32, device=loss.device)
        L.check(L.lib().mlvae_masked_reduce_fwd(
            L.ptr(loss), L.ptr(lens_f), B, T, C, L.dtype_code(loss), L.RED[reduction], L.ptr(out),
            L.ptr(L.reduce_scratch(loss.device)), L.stream_ptr()), "mlvae_masked_reduce_fwd")
        ctx.save_for_backward(lens_f)
        ctx.meta = (tuple(loss.shape), loss.dtype, reduction, B, T, C)
        return out if reduction == "batch" else out[0]

    @staticmethod
    def backward(ctx, g):
        (lens_f,) = ctx.saved_tensors
        shape, dtype, reduction, B, T, C = ctx.meta
        g = g.contiguous().float().reshape(-1)
        gl = torch.empty(shape, dtype=dtype, device=g.device)
        L.check(L.lib().mlvae_masked_reduce_bwd(
            L.ptr(g), L.ptr(lens_f), B, T, C, L.dtype_code(gl), L.RED[reduction], L.ptr(gl), L.stream_ptr()),
            "mlvae_masked_reduce_bwd")
        return gl, None, None


def masked_reduce(loss, lens, reduction: str = "mean"):
    return _MaskedReduce.apply(loss, lens, reduction)


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed, offset, offset_dev):
        L.require_cuda(x)
        x = x.contiguous()
        y = torch.empty_like(x)
        L.check(L.lib().mlvae_dropout(L.ptr(x), L.ptr(y), x.numel(), float(p), seed, offset, L.ptr(offset_dev), L.dtype_code(x),
                                      L.stream_ptr()), "mlvae_dropout")
        ctx.args = (float(p), seed, offset, offset_dev)
        return y

    @staticmethod
    def backward(ctx, dy):
        p, seed, offset, offset_dev = ctx.args
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        L.check(L.lib().mlvae_dropout(L.ptr(dy), L.ptr(dx), dy.numel(), p, seed, offset, L.ptr(offset_dev), L.dtype_code(dy),
                                      L.stream_ptr()), "mlvae_dropout")
        return dx, None, None, None, None


def dropout(x: torch.Tensor, p: float, seed: int, offset: int = 0, offset_dev=None) -> torch.Tensor:
    """Counter-based dropout (csrc/dropout.cu): y = keep ? x / (1 - p) : 0 with the Philox mask of (seed, offset
    [+ offset_dev[0], a device int64 step counter]); the backward regenerates the mask.  Stands in for the inter-layer
    dropout of nn.LSTM (modules/decoder.py:14-15)."""
    if p <= 0.0:
        return x
    return _Dropout.apply(x, p, int(seed), int(offset), offset_dev)
