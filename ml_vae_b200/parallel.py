"""Data-parallel plumbing: batch sharding and the single-bucket gradient exchange.

One process per GPU (torch.distributed, NCCL over NVLink on the box; gloo in CPU tests).
The hot path shards by utterance: rank r owns utterances
[r*B/G, (r+1)*B/G) of the global batch; the only collective on the data path is ONE
all-reduce(sum) of the flat float32 gradient bucket followed by a 1/G scale.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split; the first (global_batch % world) ranks get one extra utterance."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank: int, world: int):
    """Slice every (B, ...) tensor of a global batch to this rank's utterances."""
    b = tensors[0].shape[0]
    lo, hi = shard_bounds(b, rank, world)
    return [t[lo:hi] for t in tensors]


def all_reduce_mean_(bucket: torch.Tensor, world: int, group=None) -> torch.Tensor:
    if world > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
        bucket.mul_(1.0 / world)
    return bucket


def broadcast_(flat_params: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """Make every rank start from rank `src`'s weights (what DDP does at wrap time)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat_params, src=src, group=group)
    return flat_params


def valid_frame_count(rel_lens: torch.Tensor, T: int) -> torch.Tensor:
    """Number of frames the reference's mask keeps (utils/data_utils.py:88: ``arange(T) < lens * T`` in float32), as a
    float32 scalar on ``rel_lens``' device."""
    t = torch.arange(T, device=rel_lens.device, dtype=torch.float32)
    return (t[None, :] < (rel_lens.float()[:, None] * T)).sum().float()


def global_batch_scale(rel_lens: torch.Tensor, T: int, world: int, group=None) -> torch.Tensor:
    """Factor that turns this rank's length-masked MEAN loss into its share of the GLOBAL batch's masked mean once the gradients are
    averaged over the ranks (optional): local mean = sum_r / cnt_r, global mean = sum_all / cnt_all, and the
    average over ranks of  (world * cnt_r / cnt_all) * (sum_r / cnt_r)  is exactly sum_all / cnt_all.  One all-reduce of one float.
    With equal numbers of valid frames on every rank the factor is 1."""
    cnt = valid_frame_count(rel_lens, T)
    if world <= 1:
        return torch.ones_like(cnt)
    tot = cnt.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    return cnt * float(world) / tot
