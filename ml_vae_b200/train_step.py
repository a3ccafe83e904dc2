"""The data-parallel training step of the hot path: fused fbank -> normaliser -> VanillaVAE ->
Decoder -> masked losses -> backward -> gradient all-reduce -> Adam.

Mirrors, for the test_vanilla_vae recipe, what the reference does per batch in
  models/md_model.py:54-88    MDModel.fit_batch  (non-AMP branch)
  models/test_vanilla_vae/model.py:19-55   compute_forward / compute_objectives
  models/md_model.py:189-213  compute_and_save_losses (loss-weight lookup)
with the front-end moved into the step and one process per GPU.

B200-first choices
  * all parameters live in ONE flat float32 arena and all gradients in one flat float32
    bucket: the data-parallel exchange is a single NCCL all-reduce over NVLink on that
    bucket (no per-parameter buckets, no copies), and Adam is one fused launch over it;
  * utterances are sharded by batch across ranks; every rank reduces its own masked-mean
    loss (what per-module DDP around the reference would do), gradients are averaged;
  * non-finite loss skips the update on the device (fused-Adam found_inf), no host sync:
    the reference's check_gradients (md_model.py:82) does the same with a .item() sync.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from .parallel import all_reduce_mean_, broadcast_

KLD_N_SAMPLES = 2249          # md_model.py:199


def loss_weight(hparams: dict, loss_key: str) -> float:
    """md_model.py:192-201: x_loss -> x_weight (default 1 if absent); keys containing '_kld'
    are divided by n_samples / batch_size."""
    wkey = loss_key.replace("_loss", "_weight")
    w = hparams.get(wkey, 1)
    if "_kld" in wkey:
        w = w / (KLD_N_SAMPLES / hparams["batch_size"])
    return w


class FlatArena:
    """Re-homes every parameter of ``modules`` into one contiguous float32 buffer (and its gradient into one contiguous
    bucket) without changing names or values, plus what the fused optimiser / tensor-core kernels want next to it:
    Adam's two moment buffers, and a bf16 SHADOW of the parameters that the optimiser kernel refreshes after every update
    (``flat_bf16``; no per-layer cast kernels in the step).  Parameters a module wants stacked into one GEMM operand
    (``module.adjacent_param_groups()``) are laid out back to back; every parameter starts at a multiple of 8 elements
    (16-byte aligned bf16 / float32 rows for TMA and vector loads)."""

    ALIGN = 8

    def __init__(self, modules, peer_memory=None):
        """``peer_memory``: callable n -> peer.PeerArenaMemory (or None): the gradient / parameter / bf16 arrays are then views of
        a symmetric allocation the other ranks of the node have mapped (data-parallel optimiser step over NVLink peer memory)."""
        params, seen = [], set()
        for m in modules:
            for p in m.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        # members of an adjacency group follow the group's first member
        follow = {}
        for m in modules:
            for grp in getattr(m, "adjacent_param_groups", lambda: [])():
                for q in grp[1:]:
                    follow[id(q)] = grp[0]
        ordered, placed = [], set()
        groups = {}
        for m in modules:
            for grp in getattr(m, "adjacent_param_groups", lambda: [])():
                groups[id(grp[0])] = grp
        for p in params:
            if id(p) in placed or id(p) in follow:
                continue
            for q in groups.get(id(p), [p]):
                ordered.append(q)
                placed.add(id(q))
        self.params = ordered
        self.offsets, off = {}, 0
        for p in ordered:
            off = -(-off // self.ALIGN) * self.ALIGN
            self.offsets[id(p)] = off
            off += p.numel()
        n = -(-off // self.ALIGN) * self.ALIGN
        dev = ordered[0].device
        self.peer = peer_memory(n, dev) if peer_memory is not None else None
        if self.peer is not None:
            self.flat, self.grad, self.flat_bf16 = self.peer.flat, self.peer.grad, self.peer.flat_bf16
        else:
            self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
            self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
            self.flat_bf16 = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p in ordered:
                o, k = self.offsets[id(p)], p.numel()
                self.flat[o:o + k].copy_(p.reshape(-1))
                p.data = self.flat[o:o + k].view(p.shape)
                p.grad = self.grad[o:o + k].view(p.shape)
            self.flat_bf16.copy_(self.flat)
        self.master = nn.Parameter(self.flat, requires_grad=True)   # what a torch optimiser would see
        self.master.grad = self.grad

    def zero_grad(self):
        self.grad.zero_()

    def refresh_bf16(self):
        """Re-derive the bf16 shadow from the float32 masters (after load_state_dict / a foreign optimiser step)."""
        self.flat_bf16.copy_(self.flat)

    def all_reduce_mean(self, world_size: int, group=None, lo: int = 0, hi: int = None):
        all_reduce_mean_(self.grad[lo:hi], world_size, group)

    def offset_of(self, param) -> int:
        """Element offset of ``param`` inside the flat buffers."""
        try:
            return self.offsets[id(param)]
        except KeyError:
            raise KeyError("parameter is not in this arena") from None

    def linear_views(self, weights, biases):
        """(w_bf16 (N, K), bias_f32 (N), grad_w (N, K), grad_b (N), first master weight) over the arena for one Linear, or for several Linears with
        the same K whose weights (and biases) sit back to back (then N = sum N_i: one stacked GEMM operand).  None if the
        layout does not allow it (shapes not multiples of 8, members not adjacent)."""
        K = weights[0].shape[1]
        N = sum(w.shape[0] for w in weights)
        if K % 8 or N % 8 or any(w.shape[1] != K or w.shape[0] % 8 for w in weights):
            return None
        def run(ps):
            o0 = self.offset_of(ps[0])
            o = o0
            for p in ps:
                if self.offset_of(p) != o:
                    return None
                o += p.numel()
            return o0
        try:
            ow, ob = run(weights), run(biases)
        except KeyError:
            return None
        if ow is None or ob is None:
            return None
        return (self.flat_bf16[ow:ow + N * K].view(N, K), self.flat[ob:ob + N], self.grad[ow:ow + N * K].view(N, K), self.grad[ob:ob + N],
                weights[0])

    def broadcast(self, src: int = 0, group=None):
        broadcast_(self.flat, src, group)
        self.refresh_bf16()


class TrainStep:
    def __init__(self, fbank, normalizer, encoder, decoder, hparams: dict, lr: float = 1e-3,
                 compute_dtype=torch.bfloat16, max_grad_norm: float = 5.0, world_size: int = 1,
                 seed: int = 123456, overlap_all_reduce: bool = False, dp_mode: str = "auto",
                 defer_weight_grads: bool = True, global_batch_mean: bool = False):
        """``dp_mode`` (world_size > 1): "peer" = gradient reduce-scatter + sharded clip/Adam + parameter all-gather by the two
        kernels of csrc/dp_optim.cu over NVLink peer memory; "nccl" = one NCCL all-reduce of the flat bucket, then the full
        Adam on every rank; "auto" = peer when the node's symmetric memory can be set up (all ranks agree), else nccl."""
        self.fbank, self.normalizer, self.encoder, self.decoder = fbank, normalizer, encoder, decoder
        self.hparams = dict(hparams)
        self.dtype = compute_dtype
        self.world_size = world_size
        # exact global-batch masked means under data parallel (optional; off = every rank's own masked mean, as DDP around the reference)
        self.global_batch_mean = bool(global_batch_mean)
        self.max_grad_norm = max_grad_norm
        if dp_mode not in ("auto", "peer", "nccl"):
            raise ValueError(f"dp_mode must be 'auto', 'peer' or 'nccl', got {dp_mode!r}")
        peer_alloc = None
        if world_size > 1 and dp_mode != "nccl" and compute_dtype == torch.bfloat16 and not overlap_all_reduce:
            from . import peer as _peer
            peer_alloc = lambda n, dev: _peer.try_peer_memory(n, dev)
        self.arena = FlatArena([encoder, decoder], peer_memory=peer_alloc)
        if world_size > 1 and dp_mode == "peer" and self.arena.peer is None:
            raise RuntimeError("dp_mode='peer': the peer-memory arena could not be set up on this node")
        self.dp_peer = self.arena.peer is not None
        # optimizer: !name:torch.optim.Adam {lr: 0.001}  (models/test_vanilla_vae/model.yaml:45-47): torch's defaults, run by the
        # fused clip + Adam + zero_grad + bf16-shadow kernel pair of csrc/optim.cu over the flat arena
        self.lr, self.betas, self.eps = lr, (0.9, 0.999), 1e-8
        self.adam_state = torch.zeros(L.lib().mlvae_adam_state_bytes() // 4, dtype=torch.float32, device=self.arena.flat.device)
        if world_size > 1:
            self.arena.broadcast(0)                            # every rank starts from rank 0's weights (what DDP does at wrap time)
        for m in (encoder, decoder):
            if compute_dtype == torch.bfloat16 and hasattr(m, "bind_arena"):
                m.bind_arena(self.arena)                       # bf16 shadow weights + gradients accumulated in place
        if self.dp_peer:
            self._dp_args = L.DpAdamArgs()
            self.arena.peer.fill_args(self._dp_args)
            self._dp_args.exp_avg, self._dp_args.exp_avg_sq = self.arena.exp_avg.data_ptr(), self.arena.exp_avg_sq.data_ptr()
            self._dp_args.lr, self._dp_args.beta1, self._dp_args.beta2, self._dp_args.eps = lr, self.betas[0], self.betas[1], self.eps
            self._dp_args.max_grad_norm = float(max_grad_norm or 0.0)
        self.w_kld = loss_weight(self.hparams, "kld_loss")
        self.w_rec = loss_weight(self.hparams, "recon_loss")
        self.epoch = 0
        self.encoder.set_seed(seed)
        self.encoder.materialize_loss = False
        self.decoder.materialize_loss = False
        if hasattr(self.decoder, "direct_param_grads"):
            self.decoder.direct_param_grads = True         # parameter gradients of the LSTM land in the flat bucket directly
            if hasattr(self.decoder, "defer_weight_grads") and defer_weight_grads and not overlap_all_reduce:
                # the upper layers' weight-gradient GEMMs run on the SMs the lower layer's recurrence leaves idle (lstm.py)
                self.decoder.defer_weight_grads = True
        self.last = {}
        # device-resident step counter = Philox offset of the reparameterisation noise: a captured CUDA graph then
        # draws fresh eps on every replay
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=self.arena.flat.device)
        self.encoder.offset_dev = self.step_counter
        if hasattr(self.decoder, "dropout_offset_dev"):
            self.decoder.dropout_offset_dev = self.step_counter     # fresh inter-layer dropout mask on every graph replay
            self.decoder.dropout_seed = (seed ^ 0x5DEECE66D) & 0xFFFFFFFFFFFFFFFF
        self._graph = None
        # Data parallel, optional (overlap_all_reduce=True): the gradients of the top LSTM layer and of the heads (the tail
        # of the bucket, 72 % of its bytes for the benchmark recipe) are final when the layer below starts its backward.
        # Their all-reduce is issued on a side stream right AFTER that layer's recurrence kernel, so it overlaps the
        # layer's weight-gradient GEMMs and the encoder's backward; the head of the bucket follows after backward.
        # Bit-identical results (tests/probes/overlap_check.py), but OFF by default -- measured, same box, A/B: 6.40 vs
        # 6.39 ms per step at 2 GPUs and 6.49 vs 6.46 ms at 8 (issued BEFORE the recurrence, next to the latency-bound
        # kernel, it was worse still: 6.52 vs 6.45 and 6.71 vs 6.54 ms).  The 35 MB all-reduce costs ~0.3 ms at 8 GPUs
        # and the NCCL kernel takes about as much from the kernels it runs beside as it hides.
        self._split, self._early_done, self._ar_stream = None, False, None
        top = getattr(getattr(decoder, "rnn", None), f"weight_ih_l{getattr(decoder, 'num_layers', 1) - 1}", None)
        if overlap_all_reduce and world_size > 1 and top is not None and getattr(decoder, "num_layers", 1) > 1 and hasattr(decoder, "top_layer_grad_hook"):
            self._split = self.arena.offset_of(top)
            self._ar_stream = torch.cuda.Stream(self.arena.flat.device)
            decoder.top_layer_grad_hook = self._early_all_reduce

    def _early_all_reduce(self):
        cur = torch.cuda.current_stream()
        self._ar_stream.wait_stream(cur)                      # the tail's gradients were accumulated by earlier backward nodes
        with torch.cuda.stream(self._ar_stream):
            import torch.distributed as dist
            dist.all_reduce(self.arena.grad[self._split:], op=dist.ReduceOp.SUM)
        self._early_done = True
        return None

    # -- optimiser state (the reference registers the optimiser with its Checkpointer: models/md_model.py:50-52) ---------------
    def optimizer_state_dict(self) -> dict:
        """{'step', 'exp_avg', 'exp_avg_sq'} over the flat arena.  In the peer-memory data-parallel mode every rank updates (and so
        holds) only its shard of the two moment buffers: the shards are gathered here, so EVERY rank must call this (collective)."""
        a = self.arena
        m, v = a.exp_avg.clone(), a.exp_avg_sq.clone()
        if self.dp_peer:
            import torch.distributed as dist
            W, n = self.world_size, a.flat.numel()
            per = -(-(n // 4) // W) * 4
            for buf in (m, v):
                pad = torch.zeros(per * W, dtype=buf.dtype, device=buf.device)
                r = a.peer.rank
                lo, hi = min(n, r * per), min(n, (r + 1) * per)
                mine = torch.zeros(per, dtype=buf.dtype, device=buf.device)
                mine[:hi - lo] = buf[lo:hi]
                dist.all_gather_into_tensor(pad, mine)
                buf.copy_(pad[:n])
            step = a.peer.read_state()["step"]
        else:
            step = int(self.adam_state[0].item())
        return {"step": step, "exp_avg": m, "exp_avg_sq": v}

    def load_optimizer_state_dict(self, sd: dict):
        a = self.arena
        a.exp_avg.copy_(sd["exp_avg"].to(a.exp_avg.device))
        a.exp_avg_sq.copy_(sd["exp_avg_sq"].to(a.exp_avg_sq.device))
        if self.dp_peer:
            L.check(L.lib().mlvae_dp_set_adam_step(L.ptr(a.peer.sync), float(sd["step"]), L.stream_ptr()), "mlvae_dp_set_adam_step", kernels=0)
        else:
            self.adam_state[0] = float(sd["step"])

    # -- forward pieces -------------------------------------------------------------------
    def features(self, wav, wav_lens):
        feats, rel = self.fbank(wav, wav_lens, truncate=True, out_dtype=torch.float32)
        return self.normalizer(feats, rel, epoch=self.epoch, out_dtype=self.dtype), rel

    def losses(self, feats, rel):
        eo = self.encoder(feats, lens=rel)
        do = self.decoder(eo["sampled_h"], feats, lens=rel)
        kld, rec = eo["kld_loss"], do["recon_loss"]
        loss = self.w_kld * kld + self.w_rec * rec
        if self.global_batch_mean and self.world_size > 1:
            from .parallel import global_batch_scale
            loss = loss * global_batch_scale(rel, feats.shape[1], self.world_size)      # one all-reduce of one float
        return loss, kld, rec

    # -- one training step ----------------------------------------------------------------
    def step_from_features(self, feats, rel):
        loss, kld, rec = self.losses(feats, rel)
        self._early_done = False
        loss.backward()
        from . import lstm as _lstm
        _lstm.flush_deferred()                                  # normally empty: the lowest LSTM layer's backward ran what was queued
        lossf = loss.detach().float().reshape(1)
        a = self.arena
        if self.dp_peer:
            # reduce-scatter of the gradients, check_gradients + Adam on this rank's shard, all-gather of the parameters and their bf16
            # shadow, zero_grad: two launches over the peers' arenas (csrc/dp_optim.cu)
            args = self._dp_args
            args.loss = L.ptr(lossf)
            self._dp_keep = lossf
            L.check(L.lib().mlvae_dp_adam_step(args, L.stream_ptr()), "mlvae_dp_adam_step", kernels=2)
            self.step_counter += 1
            self.last = {"loss": loss.detach(), "kld_loss": kld.detach(), "recon_loss": rec.detach()}
            return self.last["loss"]
        # gradient SUM over the ranks (the 1 / world_size of the mean is applied inside the optimiser kernel)
        if self.world_size > 1:
            import torch.distributed as dist
            if self._early_done:
                dist.all_reduce(self.arena.grad[:self._split], op=dist.ReduceOp.SUM)
                torch.cuda.current_stream().wait_stream(self._ar_stream)
            else:
                dist.all_reduce(self.arena.grad, op=dist.ReduceOp.SUM)
        # check_gradients [SB-recall] (non-finite loss -> skip the update; clip the global norm), Adam, zero_grad, bf16 shadow
        L.check(L.lib().mlvae_adam_clip_step(L.ptr(a.flat), L.ptr(a.grad), L.ptr(a.exp_avg), L.ptr(a.exp_avg_sq), L.ptr(a.flat_bf16), a.flat.numel(),
                                             1.0 / self.world_size, self.lr, self.betas[0], self.betas[1], self.eps, float(self.max_grad_norm or 0.0),
                                             L.ptr(self.adam_state), L.ptr(lossf), L.stream_ptr()), "mlvae_adam_clip_step", kernels=2)
        self.step_counter += 1
        self.last = {"loss": loss.detach(), "kld_loss": kld.detach(), "recon_loss": rec.detach()}
        return self.last["loss"]

    def _step_eager(self, wav, wav_lens):
        feats, rel = self.features(wav, wav_lens)
        return self.step_from_features(feats, rel)

    def step(self, wav, wav_lens):
        """wav (B, N) float32 on the device, wav_lens (B,) absolute samples or relative lengths."""
        if self._graph is not None and wav.shape == self._g_wav.shape:
            self._g_wav.copy_(wav, non_blocking=True)
            self._g_lens.copy_(wav_lens.to(self._g_lens.dtype), non_blocking=True)
            self._graph.replay()
            return self.last["loss"]
        return self._step_eager(wav, wav_lens)

    # -- host-fed stepping with the H2D copy of the next batch overlapped --------------------------
    def submit_host_batch(self, host_wav: torch.Tensor):
        """Start copying a (pinned) host batch to the device on a side stream; consumed by the next
        ``step_submitted``.  Two staging buffers: batch i+1 streams in over PCIe while step i computes."""
        if getattr(self, "_copy_stream", None) is None:
            dev = self.arena.flat.device
            self._copy_stream = torch.cuda.Stream(dev)
            self._stage = [torch.empty(host_wav.shape, dtype=torch.float32, device=dev) for _ in range(2)]
            self._ready = [torch.cuda.Event() for _ in range(2)]
            self._free = [torch.cuda.Event() for _ in range(2)]
            for e in self._free:
                e.record()
            self._pf, self._pending = 0, []
        k = self._pf
        self._copy_stream.wait_event(self._free[k])            # the step that last read this buffer has finished
        with torch.cuda.stream(self._copy_stream):
            self._stage[k].copy_(host_wav, non_blocking=True)
            self._ready[k].record(self._copy_stream)
        self._pending.append(k)
        self._pf ^= 1

    def step_submitted(self, wav_lens):
        k = self._pending.pop(0)
        torch.cuda.current_stream().wait_event(self._ready[k])
        loss = self.step(self._stage[k], wav_lens)
        self._free[k].record()
        return loss

    # -- CUDA graph ---------------------------------------------------------------------------
    def capture(self, wav, wav_lens, warmup: int = 3) -> bool:
        """Capture the whole step (front-end, forward, backward, all-reduce, clip, Adam) for this input shape in
        one CUDA graph; later ``step`` calls with the same shape replay it.  All per-step state that used to live
        on the host (Philox offset, normaliser count, non-finite skip) is device resident, so replays are real
        training steps.  ``warmup`` eager steps run first (they are training steps too).  Returns False and stays
        eager if capture is refused (e.g. by the communicator)."""
        self._g_wav = wav.clone()
        self._g_lens = wav_lens.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._step_eager(self._g_wav, self._g_lens)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._step_eager(self._g_wav, self._g_lens)
            self._graph = graph
            return True
        except Exception as exc:                       # stay correct: fall back to eager stepping
            import warnings
            warnings.warn(f"CUDA-graph capture of the training step failed, staying eager: {exc}")
            self._graph = None
            torch.cuda.synchronize()
            return False
