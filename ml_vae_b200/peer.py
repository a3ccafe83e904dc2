"""Peer-memory plumbing of the data-parallel optimiser step (csrc/dp_optim.cu, ``mlvae_dp_adam_step``).

The kernels load the other ranks' gradient arenas and store into their parameter arenas over NVLink, so those arrays
must live in memory every rank of the node has mapped.  This module only allocates and exchanges mappings (torch's
symmetric-memory allocator: CUDA VMM allocations, handles exchanged through the process group's store, a multicast
(NVLS) mapping where the fabric offers one); the reduction, the optimiser and the inter-rank barriers are the kernels'.

One symmetric allocation per rank, laid out as  [ gradients f32 n | parameters f32 n | bf16 shadow n | sync block ],
so a single rendezvous yields every peer pointer (and ONE multicast mapping covers gradients, parameters and shadow).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L

MAX_WORLD = 8


class PeerArenaMemory:
    """The three peer-visible arrays of a FlatArena with ``n`` elements plus the sync block of ``mlvae_dp_adam_step``."""

    def __init__(self, n: int, device: torch.device, group=None, multicast=None):
        """``multicast``: use the NVLS mapping (multimem.ld_reduce / multimem.st) when the fabric offers one; None = only for more
        than two ranks (measured on B200 x8 for the 34.6 MB benchmark arena: 156 vs 180 us per step at 8 ranks, but 169 vs
        109 us at 2 ranks, where every multimem access crosses the switch twice for no saving)."""
        import torch.distributed._symmetric_memory as symm

        if n % 8:
            raise ValueError("arena size must be a multiple of 8 elements")
        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not 1 <= self.world <= MAX_WORLD:
            raise ValueError(f"peer-memory optimiser step: 1..{MAX_WORLD} ranks of one node, got {self.world}")
        self.n = n
        sync_bytes = L.lib().mlvae_dp_sync_bytes()
        self.off_grad, self.off_param, self.off_bf16 = 0, 4 * n, 8 * n
        self.off_sync = -(-10 * n // 256) * 256
        total = self.off_sync + -(-sync_bytes // 256) * 256
        self.buf = symm.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric memory rendezvous returned unexpected peer pointers")
        self.peer_base = ptrs
        if multicast is None:
            multicast = self.world > 2
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0) if multicast else 0
        self.multicast_base = mc
        self.grad = self.buf[self.off_grad:self.off_grad + 4 * n].view(torch.float32)
        self.flat = self.buf[self.off_param:self.off_param + 4 * n].view(torch.float32)
        self.flat_bf16 = self.buf[self.off_bf16:self.off_bf16 + 2 * n].view(torch.bfloat16)
        self.sync = self.buf[self.off_sync:self.off_sync + sync_bytes]
        torch.cuda.synchronize(device)
        self.handle.barrier()                                   # every rank's block is zeroed before anyone signals into it

    def fill_args(self, a: "L.DpAdamArgs"):
        a.world, a.rank = self.world, self.rank
        for r in range(self.world):
            base = self.peer_base[r]
            a.grads[r] = base + self.off_grad
            a.params[r] = base + self.off_param
            a.params_bf16[r] = base + self.off_bf16
            a.sync[r] = base + self.off_sync
        if self.multicast_base:
            a.mc_grads = self.multicast_base + self.off_grad
            a.mc_params = self.multicast_base + self.off_param
            a.mc_params_bf16 = self.multicast_base + self.off_bf16
        a.n = self.n

    def read_state(self) -> dict:
        out = (C.c_float * 5)()
        L.check(L.lib().mlvae_dp_read_state(L.ptr(self.sync), C.byref(out), L.stream_ptr()), "mlvae_dp_read_state", kernels=0)
        return {"epoch": int(out[0]), "step": int(out[1]), "grad_norm": float(out[2]), "clip_coef": float(out[3]), "error": int(out[4])}


def try_peer_memory(n: int, device: torch.device, group=None, multicast=None):
    """PeerArenaMemory on every rank, or None on every rank (the ranks agree through one all-reduce): the caller then keeps
    the NCCL all-reduce path."""
    mem, err = None, None
    try:
        mem = PeerArenaMemory(n, device, group, multicast)
    except Exception as exc:                                   # allocator / fabric not available on this box
        err = exc
    ok = torch.tensor([1 if mem is not None else 0], dtype=torch.int32, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        if err is not None:
            import warnings
            warnings.warn(f"peer-memory optimiser step unavailable, using the NCCL all-reduce: {type(err).__name__}: {err}")
        return None
    return mem
