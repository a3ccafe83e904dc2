"""Bidirectional LSTM layer on the persistent tcgen05 kernels (csrc/lstm.cu).

Replaces, for bf16 activations, the cuDNN path behind ``nn.LSTM(..., bidirectional=True,
batch_first=True)`` that the reference's Decoder uses (modules/decoder.py:14-15,22).  The big,
time-parallel GEMMs (input projection, dX, dW_ih, dW_hh) run on the TMA / tcgen05 GEMM of
csrc/gemm.cu; the sequential part -- T dependent steps per direction -- is ONE cooperative kernel
per pass instead of 2 x T cuDNN launches.  Gate order and parameter layout are torch's (i, f, g, o).
"""
from __future__ import annotations

import torch

from . import _lib as L

_scratch = {}
PROBE = None      # bench.py sets this to a list; every recurrence launch then appends (tag, start_event, end_event)

# Weight-gradient GEMMs of an UPPER layer that its backward left for the next recurrence launch to hide.  The persistent
# recurrence is a cooperative launch of 2 * (H / 32) * ceil(B / 16) CTAs, one per SM (128 of 148 at the benchmark shape, 32 at
# configs[3]) that spends its time waiting on the inter-CTA exchange; the SMs it does not use are idle for ~1 ms.  A layer whose
# ``defer_weight_grads`` is set therefore only QUEUES its dW GEMMs; the backward of the layer below launches them on a side
# stream, confined to the idle SMs (gemm(max_ctas=...)), right before its own recurrence, and joins the streams after it.
_DEFERRED = []
_SIDE = {}


def _side_stream(device):
    key = torch.device(device).index
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device)
    return _SIDE[key]


def idle_sms(B: int, hidden: int, device) -> int:
    """SMs the recurrence launch for (B rows, hidden) leaves idle."""
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    rows = min(B, max_rows_per_launch(hidden, device))
    return max(0, sms - 2 * (hidden // 32) * ((rows + 15) // 16))


def flush_deferred():
    """Run whatever weight-gradient GEMMs are still queued on the current stream (no recurrence left to hide them under)."""
    work = _DEFERRED[:]
    _DEFERRED.clear()
    for w in work:
        w(0)


def max_rows_per_launch(hidden: int, device) -> int:
    """Batch rows one cooperative launch can take: every (direction, 16-row slice, 32-unit slice) CTA must be
    co-resident, one per SM.  Larger batches are walked in chunks of this many rows (the recurrences of different rows
    are independent)."""
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    return max(16, (sms // (2 * (hidden // 32))) * 16)


def supported(x: torch.Tensor, hidden: int) -> bool:
    return x.is_cuda and x.dtype == torch.bfloat16 and hidden % 32 == 0 and 32 <= hidden <= 512 \
        and L.lib().mlvae_lstm_scratch_bytes(min(x.shape[0], max_rows_per_launch(hidden, x.device)), hidden) > 0


def _get_scratch(B, H, device):
    need = L.lib().mlvae_lstm_scratch_bytes(B, H)
    if need == 0:
        raise L.MlvaeError(f"persistent LSTM does not support batch {B} x hidden {H}: {L.lib().mlvae_last_error().decode()}")
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


_perm_cache = {}


def _gate_perm(H: int, device):
    """Row permutation torch order (dir, gate, unit) -> kernel order (dir, unit, gate) of the stacked (8H, .) matrices,
    and its inverse.  The recurrence kernels keep the four gates of a unit adjacent (one 8-byte access per lane)."""
    key = (H, device)
    if key not in _perm_cache:
        one = torch.arange(4 * H, device=device).view(4, H).t().reshape(-1)          # new (unit*4+gate) -> old (gate*H+unit)
        perm = torch.cat([one, one + 4 * H])
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(8 * H, device=device)
        _perm_cache[key] = (perm, inv)
    return _perm_cache[key]


def _probe_start():
    if PROBE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _probe_end(tag, start):
    if start is not None:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        PROBE.append((tag, start, e))


def _mm_f32(a, b):
    """a @ b with bf16 operands and a float32 result straight out of the GEMM (no bf16 rounding + cast kernel)."""
    return torch.mm(a, b, out_dtype=torch.float32)


def _ptr_array(tensors):
    import ctypes as C
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _packable(masters, In: int, H: int) -> bool:
    return In % 4 == 0 and H % 4 == 0 and all(m.dtype == torch.float32 and m.is_contiguous() and m.data_ptr() % 16 == 0
                                              for m in masters)


def _gemm_ok(In: int, H: int, x: torch.Tensor) -> bool:
    """Shapes the TMA GEMM takes: 16-byte aligned rows of every operand."""
    return In % 8 == 0 and H % 8 == 0 and x.data_ptr() % 16 == 0


class _BiLSTMLayer(torch.autograd.Function):
    """One bidirectional layer.  Takes torch's eight per-direction float32 master parameters directly and hands their
    gradients back in float32, so autograd needs no cat / cast / add nodes (and their kernels) around the layer.

    Every matrix product of the layer runs on the TMA / tcgen05 GEMM of csrc/gemm.cu (``gemm.gemm``): the input projection
    (+ float32 bias in the epilogue), the input gradient (+ the dropout mask of the layer's input, if any, in the epilogue) and
    the weight gradients, which the GEMM accumulates straight into torch-ordered float32 tensors -- the parameters' own
    ``.grad`` buffers when ``direct_grads`` -- with dW_hh as ONE batched GEMM over row-shifted views."""

    @staticmethod
    def forward(ctx, x, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, training, direct_grads, after_recurrence,
                input_dropout, defer_weight_grads=False):
        """x (B,T,In) bf16 -> y (B,T,2H) bf16.  ``input_dropout`` = (p, seed, offset, offset_dev) or None: the layer's input is
        dropout(x) with the counter-based mask of csrc/dropout.cu (the inter-layer dropout of nn.LSTM, decoder.py:14-15)."""
        from .gemm import gemm
        B, T, In = x.shape
        H = w_hh_f.shape[1]
        bf = torch.bfloat16
        if input_dropout is not None:
            p_drop, d_seed, d_off, d_dev = input_dropout
            xd = torch.empty_like(x)
            L.check(L.lib().mlvae_dropout(L.ptr(x), L.ptr(xd), x.numel(), float(p_drop), d_seed, d_off, L.ptr(d_dev), L.BF16, L.stream_ptr()),
                    "mlvae_dropout")
            x = xd
        x2 = x.reshape(B * T, In)
        masters = (w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r)
        if _packable(masters, In, H):
            # one launch: cast + stack + permute (csrc/lstm_pack.cu)
            w_ih_p = torch.empty(8 * H, In, dtype=bf, device=x.device)       # rows in (dir, unit, gate) order
            w_hh = torch.empty(2, 4 * H, H, dtype=bf, device=x.device)
            bias_p = torch.empty(8 * H, dtype=torch.float32, device=x.device)
            L.check(L.lib().mlvae_lstm_pack_weights(_ptr_array(masters), In, H, L.ptr(w_ih_p), L.ptr(w_hh), L.ptr(bias_p),
                                                    L.stream_ptr()), "mlvae_lstm_pack_weights")
        else:
            perm, _ = _gate_perm(H, x.device)
            w_ih_p = torch.cat([w_ih_f, w_ih_r], 0).to(bf)[perm].contiguous()
            w_hh = torch.stack([w_hh_f, w_hh_r], 0).to(bf)
            bias_p = torch.cat([b_ih_f + b_hh_f, b_ih_r + b_hh_r], 0).float()[perm].contiguous()
        ctx.masters = masters if (training and direct_grads) else None
        ctx.after_recurrence = after_recurrence
        ctx.input_dropout = input_dropout
        ctx.defer_weight_grads = bool(defer_weight_grads)
        ctx.use_gemm = _gemm_ok(In, H, x2)
        if ctx.use_gemm:
            P = torch.empty(B, T, 2, 4 * H, dtype=bf, device=x.device)
            gemm(x2, w_ih_p, P, B * T, 8 * H, In, lda=In, ldb=In, ldd=8 * H, bias=bias_p)
        else:
            P = torch.addmm(bias_p.to(bf), x2, w_ih_p.t()).view(B, T, 2, 4 * H)  # odd shapes: library GEMM
        y = torch.empty(B, T, 2 * H, dtype=bf, device=x.device)
        c = torch.empty(B, T, 2 * H, dtype=torch.float32, device=x.device) if training else None
        ev = _probe_start()
        rows = max_rows_per_launch(H, x.device)
        for b0 in range(0, B, rows):
            b1 = min(B, b0 + rows)
            L.check(L.lib().mlvae_lstm_fwd(L.ptr(P[b0:b1]), L.ptr(w_hh), L.ptr(y[b0:b1]), L.ptr(c[b0:b1]) if training else None,
                                           b1 - b0, T, H, int(training), L.ptr(_get_scratch(b1 - b0, H, x.device)), L.stream_ptr()),
                    "mlvae_lstm_fwd")
        _probe_end("lstm_fwd", ev)
        if training:
            ctx.save_for_backward(x, w_ih_p, w_hh, P, c, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        from .gemm import gemm
        x, w_ih_p, w_hh, gates, c, y = ctx.saved_tensors
        if getattr(ctx, "consumed", False):
            # the backward kernel overwrites the saved gates in place with the pre-activation gradients (raw pointers, so
            # autograd's version counter cannot see it): a second backward over the same graph would read garbage
            raise RuntimeError("bilstm_layer: backward called twice over the same forward (retain_graph / double use); "
                               "the saved gate buffer is consumed by the first backward -- run the forward again")
        ctx.consumed = True
        B, T, In = x.shape
        H = w_hh.shape[2]
        H4 = 4 * H
        dev = x.device
        dy = dy.contiguous().to(torch.bfloat16)
        ev = _probe_start()
        rows = max_rows_per_launch(H, dev)
        n_slices = sum((min(B, b0 + rows) - b0 + 15) // 16 for b0 in range(0, B, rows))
        db_part = torch.empty(n_slices, 8 * H, dtype=torch.float32, device=dev)
        part0 = 0
        # weight-gradient GEMMs the layer above queued: on a side stream, on the SMs this recurrence leaves idle
        beside, side = _DEFERRED[:], None
        _DEFERRED.clear()
        if beside:
            main, side = torch.cuda.current_stream(dev), _side_stream(dev)
            side.wait_stream(main)
            free = idle_sms(B, H, dev)
            with torch.cuda.stream(side):
                for w in beside:
                    w(free)
        for b0 in range(0, B, rows):
            b1 = min(B, b0 + rows)
            L.check(L.lib().mlvae_lstm_bwd(L.ptr(gates[b0:b1]), L.ptr(c[b0:b1]), L.ptr(dy[b0:b1]), L.ptr(w_hh), L.ptr(db_part[part0:]),
                                           b1 - b0, T, H, L.ptr(_get_scratch(b1 - b0, H, dev)), L.stream_ptr()), "mlvae_lstm_bwd")
            part0 += (b1 - b0 + 15) // 16
        _probe_end("lstm_bwd", ev)
        if beside:
            torch.cuda.current_stream(dev).wait_stream(side)               # their operands are released only after this join
            del beside
        if ctx.after_recurrence is not None:
            ctx.after_recurrence()                                         # e.g. start the all-reduce of the layers above
        dA = gates                                                         # now pre-activation gradients (B,T,2,H,4)
        dA2 = dA.view(B * T, 8 * H)                                        # columns in (dir, unit, gate) order
        x2 = x.reshape(B * T, In)
        y2 = y.view(B * T, 2 * H)
        masters = ctx.masters
        direct = masters is not None and all(m.grad is not None and m.grad.dtype == torch.float32 and m.grad.is_contiguous()
                                             and m.grad.data_ptr() % 16 == 0 for m in masters)
        if not ctx.use_gemm:
            return _BiLSTMLayer._backward_library(ctx, dA, dA2, x2, y, y2, w_ih_p, db_part, masters if direct else None, B, T, In, H)

        # ---- input gradient: dx = dA W_ih (W_ih row-major (8H, In): MN-major B), the input's dropout mask in the epilogue ----
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, T, In, dtype=torch.bfloat16, device=dev)
            kw = {}
            if ctx.input_dropout is not None:
                p_drop, d_seed, d_off, d_dev = ctx.input_dropout
                kw = dict(drop_p=p_drop, drop_seed=d_seed, drop_offset=d_off, drop_offset_dev=d_dev)
            gemm(dA2, w_ih_p, dx, B * T, In, 8 * H, lda=8 * H, ldb=In, ldd=In, b_mn=True, **kw)
        # ---- weight gradients, float32, torch (gate, unit) row order, accumulated where they live ----
        if direct:
            g_ih, g_hh = [masters[0].grad, masters[4].grad], [masters[1].grad, masters[5].grad]
            g_b = [masters[2].grad, masters[3].grad, masters[6].grad, masters[7].grad]
        else:
            g_ih = [torch.empty(H4, In, dtype=torch.float32, device=dev) for _ in range(2)]
            g_hh = [torch.zeros(H4, H, dtype=torch.float32, device=dev) for _ in range(2)]
            g_b = [torch.zeros(H4, dtype=torch.float32, device=dev) for _ in range(4)]
        # dW_ih[d] = dA[:, d]^T x: both operands MN-major; few output tiles for a narrow input -> split the B*T reduction
        # With ``defer_weight_grads`` (and gradients accumulated in place) the products are only queued: the backward of the layer
        # below runs them beside its recurrence on the SMs that launch leaves idle -- when there are enough of them: with 20 idle SMs
        # (configs[1]) the GEMM's L2 traffic slows the latency-bound recurrence by 0.09 ms while hiding 0.13 ms, not worth the
        # disturbed kernel; with 116 idle SMs (configs[3]) both products disappear under the recurrence (15.3 -> 14.9 ms per step).
        free = idle_sms(B, H, dev) if (ctx.defer_weight_grads and direct) else 0

        def dw_ih(max_ctas=0):
            tiles = 2 * ((H4 + 127) // 128) * ((In + 255) // 256)
            split = 1 if tiles >= 96 else max(1, min(8, 128 // tiles))
            gemm([dA2[:, :H4], dA2[:, H4:]], [x2, x2], g_ih, H4, In, B * T, lda=8 * H, ldb=In, ldd=In, a_mn=True, b_mn=True, out_f32=True,
                 accumulate=direct, row_perm_H=H, split_k=split, max_ctas=max_ctas)

        def dw_hh(max_ctas=0):
            # dW_hh[d] = sum_b sum_t dA[b,t,d]^T h_prev[b,t,d], h_prev the previous step IN THAT DIRECTION'S ORDER (forward: y[b,t-1,:H];
            # reverse: y[b,t+1,H:]): a batched reduction over row-shifted views, T-1 rows per utterance (TMA zero-fills past them)
            tiles = 2 * ((H4 + 127) // 128) * ((H + 255) // 256)
            split = 1 if tiles >= 96 else max(1, min(4, 128 // tiles))
            gemm([dA2[1:, :H4], dA2[:, H4:]], [y2[:, :H], y2[1:, H:]], g_hh, H4, H, T - 1, lda=8 * H, ldb=2 * H, ldd=H, a_mn=True, b_mn=True,
                 kbatches=B, a_batch_stride=T * 8 * H, b_batch_stride=T * 2 * H, out_f32=True, accumulate=direct, row_perm_H=H, split_k=split,
                 max_ctas=max_ctas)

        if free >= 64:
            _DEFERRED.append(dw_ih)
        else:
            dw_ih()
        if T > 1:
            if free >= 32:
                _DEFERRED.append(dw_hh)
            else:
                dw_hh()
        L.check(L.lib().mlvae_lstm_bias_grads(L.ptr(db_part), n_slices, H, L.ptr(g_b[0]), L.ptr(g_b[1]), L.ptr(g_b[2]), L.ptr(g_b[3]), L.stream_ptr()),
                "mlvae_lstm_bias_grads")
        if direct:
            return (dx,) + (None,) * 13
        return dx, g_ih[0], g_hh[0], g_b[0], g_b[1], g_ih[1], g_hh[1], g_b[2], g_b[3], None, None, None, None, None

    @staticmethod
    def _backward_library(ctx, dA, dA2, x2, y, y2, w_ih_p, db_part, masters, B, T, In, H):
        """Input sizes the TMA GEMM does not take (In or H not a multiple of 8): the same products as library GEMMs."""
        H4 = 4 * H
        dev = x2.device
        _, inv = _gate_perm(H, dev)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = (dA2 @ w_ih_p).view(B, T, In)
            if ctx.input_dropout is not None:
                p_drop, d_seed, d_off, d_dev = ctx.input_dropout
                L.check(L.lib().mlvae_dropout(L.ptr(dx), L.ptr(dx), dx.numel(), float(p_drop), d_seed, d_off, L.ptr(d_dev), L.BF16, L.stream_ptr()),
                        "mlvae_dropout")
        dw_ih = _mm_f32(dA2.t(), x2)[inv]
        db = db_part.sum(0)
        if T > 1:
            g0 = _mm_f32(dA2[1:, :H4].t(), y2[:-1, :H])
            g1 = _mm_f32(dA2[:-1, H4:].t(), y2[1:, H:])
            if B > 1:
                g0 -= _mm_f32(dA[1:, 0, 0].t(), y[:-1, T - 1, :H])
                g1 -= _mm_f32(dA[:-1, T - 1, 1].t(), y[1:, 0, H:])
            dw_hh_f, dw_hh_r = g0[inv[:H4]], g1[inv[:H4]]
        else:
            dw_hh_f = torch.zeros(H4, H, dtype=torch.float32, device=dev)
            dw_hh_r = torch.zeros_like(dw_hh_f)
        grads = [dw_ih[:H4], dw_hh_f, db[:H4], db[:H4], dw_ih[H4:], dw_hh_r, db[H4:], db[H4:]]
        if masters is not None:
            for m, g in zip(masters, grads):
                m.grad.add_(g)
            return (dx,) + (None,) * 13
        return (dx, *grads, None, None, None, None, None)


def bilstm_layer(x, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, training: bool, direct_grads: bool = False,
                 after_recurrence=None, input_dropout=None, defer_weight_grads: bool = False):
    """One bidirectional layer with torch's per-direction parameters (float32 masters; cast inside the layer).
    ``direct_grads``: accumulate the parameter gradients straight into the parameters' existing float32 ``.grad``
    buffers (valid under ``loss.backward()``; the training step that owns a flat gradient bucket turns it on).
    ``after_recurrence``: callable run in backward right after the recurrence kernel is enqueued, before the layer's
    weight-gradient GEMMs.  ``input_dropout`` = (p, seed, offset, offset_dev): the layer consumes dropout(x) (counter-based
    mask, csrc/dropout.cu); the mask of the backward pass is applied inside the input-gradient GEMM.
    ``defer_weight_grads`` (needs ``direct_grads``): backward only queues the layer's weight-gradient GEMMs; the backward of the
    NEXT bilstm_layer runs them beside its recurrence on the idle SMs, ``flush_deferred()`` runs whatever is left -- the caller
    guarantees one of the two happens before the gradients are read (Decoder: every layer but the lowest; TrainStep flushes)."""
    return _BiLSTMLayer.apply(x.contiguous(), w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, training,
                              direct_grads, after_recurrence, input_dropout, defer_weight_grads)


# ------------------------------------------------------------------------------------------------------------------------------
# Forward-only layer: nn.LSTM(batch_first=True) without `bidirectional` -- the main RNN of the MD_VAE* recipes
# (models/MD_VAE/model.yaml:78-83), modules/boundary_detector.py:19, modules/phoneme_recognizer.py:13.  Same two recurrence kernels
# (their one-direction instantiation, mlvae_lstm_fwd_dirs / _bwd_dirs with ndir = 1) and the same GEMM for the time-parallel
# products; the parameter plumbing (cast / gate-row permutation of four small tensors) is left to torch here.
# ------------------------------------------------------------------------------------------------------------------------------
def max_rows_per_launch_dirs(hidden: int, device, ndir: int) -> int:
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    return max(16, (sms // (ndir * (hidden // 32))) * 16)


def supported_uni(x: torch.Tensor, hidden: int) -> bool:
    return x.is_cuda and x.dtype == torch.bfloat16 and hidden % 32 == 0 and 32 <= hidden <= 512 and x.shape[-1] % 8 == 0 \
        and L.lib().mlvae_lstm_scratch_bytes_dirs(min(x.shape[0], max_rows_per_launch_dirs(hidden, x.device, 1)), hidden, 1) > 0


def _scratch_dirs(B, H, ndir, device):
    need = L.lib().mlvae_lstm_scratch_bytes_dirs(B, H, ndir)
    if need == 0:
        raise L.MlvaeError(f"persistent LSTM does not support batch {B} x hidden {H}: {L.lib().mlvae_last_error().decode()}")
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


class _UniLSTMLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, training, input_dropout):
        from .gemm import gemm
        B, T, In = x.shape
        H = w_hh.shape[1]
        H4 = 4 * H
        bf, dev = torch.bfloat16, x.device
        if not _gemm_ok(In, H, x):
            raise NotImplementedError(f"lstm_layer: input size {In} / hidden size {H} must be multiples of 8 (16-byte rows for TMA)")
        if input_dropout is not None:
            p_drop, d_seed, d_off, d_dev = input_dropout
            xd = torch.empty_like(x)
            L.check(L.lib().mlvae_dropout(L.ptr(x), L.ptr(xd), x.numel(), float(p_drop), d_seed, d_off, L.ptr(d_dev), L.BF16, L.stream_ptr()),
                    "mlvae_dropout")
            x = xd
        x2 = x.reshape(B * T, In)
        perm = _gate_perm(H, dev)[0][:H4]                                  # kernel row unit*4+gate <- torch row gate*H+unit
        w_ih_p = w_ih.to(bf)[perm].contiguous()
        w_hh_b = w_hh.to(bf).contiguous().view(1, H4, H)                   # torch row order (the kernels index it by gate*H+unit)
        bias_p = (b_ih + b_hh).float()[perm].contiguous()
        P = torch.empty(B, T, 1, H4, dtype=bf, device=dev)
        gemm(x2, w_ih_p, P, B * T, H4, In, lda=In, ldb=In, ldd=H4, bias=bias_p)
        y = torch.empty(B, T, H, dtype=bf, device=dev)
        c = torch.empty(B, T, H, dtype=torch.float32, device=dev) if training else None
        rows = max_rows_per_launch_dirs(H, dev, 1)
        for b0 in range(0, B, rows):
            b1 = min(B, b0 + rows)
            L.check(L.lib().mlvae_lstm_fwd_dirs(L.ptr(P[b0:b1]), L.ptr(w_hh_b), L.ptr(y[b0:b1]), L.ptr(c[b0:b1]) if training else None,
                                                b1 - b0, T, H, 1, int(training), L.ptr(_scratch_dirs(b1 - b0, H, 1, dev)), L.stream_ptr()),
                    "mlvae_lstm_fwd_dirs")
        ctx.input_dropout = input_dropout
        ctx.param_dtype = w_ih.dtype
        if training:
            ctx.save_for_backward(x, w_ih_p, w_hh_b, P, c, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        from .gemm import gemm
        x, w_ih_p, w_hh_b, gates, c, y = ctx.saved_tensors
        if getattr(ctx, "consumed", False):
            raise RuntimeError("lstm_layer: backward called twice over the same forward; the saved gate buffer is consumed by the "
                               "first backward -- run the forward again")
        ctx.consumed = True
        B, T, In = x.shape
        H = w_hh_b.shape[2]
        H4, dev = 4 * H, x.device
        dy = dy.contiguous().to(torch.bfloat16)
        rows = max_rows_per_launch_dirs(H, dev, 1)
        n_slices = sum((min(B, b0 + rows) - b0 + 15) // 16 for b0 in range(0, B, rows))
        db_part = torch.empty(n_slices, H4, dtype=torch.float32, device=dev)
        part0 = 0
        for b0 in range(0, B, rows):
            b1 = min(B, b0 + rows)
            L.check(L.lib().mlvae_lstm_bwd_dirs(L.ptr(gates[b0:b1]), L.ptr(c[b0:b1]), L.ptr(dy[b0:b1]), L.ptr(w_hh_b), L.ptr(db_part[part0:]),
                                                b1 - b0, T, H, 1, L.ptr(_scratch_dirs(b1 - b0, H, 1, dev)), L.stream_ptr()), "mlvae_lstm_bwd_dirs")
            part0 += (b1 - b0 + 15) // 16
        dA2 = gates.view(B * T, H4)                                         # pre-activation gradients, columns in (unit, gate) order
        x2, y2 = x.reshape(B * T, In), y.view(B * T, H)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B, T, In, dtype=torch.bfloat16, device=dev)
            kw = {}
            if ctx.input_dropout is not None:
                p_drop, d_seed, d_off, d_dev = ctx.input_dropout
                kw = dict(drop_p=p_drop, drop_seed=d_seed, drop_offset=d_off, drop_offset_dev=d_dev)
            gemm(dA2, w_ih_p, dx, B * T, In, H4, lda=H4, ldb=In, ldd=In, b_mn=True, **kw)
        g_ih = torch.empty(H4, In, dtype=torch.float32, device=dev)
        g_hh = torch.zeros(H4, H, dtype=torch.float32, device=dev)
        tiles = ((H4 + 127) // 128) * ((In + 255) // 256)
        gemm([dA2], [x2], [g_ih], H4, In, B * T, lda=H4, ldb=In, ldd=In, a_mn=True, b_mn=True, out_f32=True, row_perm_H=H,
             split_k=1 if tiles >= 96 else max(1, min(8, 128 // tiles)))
        if T > 1:
            # dW_hh = sum_b sum_t dA[b, t+1]^T h[b, t]: one batched GEMM over row-shifted views (TMA zero-fills past T-1 rows)
            tiles = ((H4 + 127) // 128) * ((H + 255) // 256)
            gemm([dA2[1:]], [y2], [g_hh], H4, H, T - 1, lda=H4, ldb=H, ldd=H, a_mn=True, b_mn=True, kbatches=B, a_batch_stride=T * H4,
                 b_batch_stride=T * H, out_f32=True, row_perm_H=H, split_k=1 if tiles >= 96 else max(1, min(4, 128 // tiles)))
        db = db_part.sum(0)                                                 # torch gate order already; both biases get it
        pd = ctx.param_dtype                                                # float32 masters (or a module converted to bf16)
        return dx, g_ih.to(pd), g_hh.to(pd), db.to(pd), db.to(pd).clone(), None, None


def lstm_layer(x, w_ih, w_hh, b_ih, b_hh, training: bool, input_dropout=None):
    """One forward-direction LSTM layer with torch's parameters (float32 masters): x (B, T, In) bf16 -> (B, T, H) bf16."""
    return _UniLSTMLayer.apply(x.contiguous(), w_ih, w_hh, b_ih, b_hh, training, input_dropout)
