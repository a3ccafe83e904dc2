"""CPU, world_size 2, gloo: the N>1 path of the training step -- batch sharding and the
single flat-bucket gradient all-reduce -- without any kernel call."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ml_vae_b200.parallel import shard_bounds


def test_shard_bounds_cover_batch_exactly():
    for B in (1, 7, 64, 65, 512):
        for G in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ml_vae_b200.parallel import shard_batch
    from ml_vae_b200.train_step import FlatArena
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    if rank == 1:                                   # deliberately different start: broadcast must fix it
        with torch.no_grad():
            for p in model.parameters():
                p.add_(1.0)
    arena = FlatArena([model])
    arena.broadcast(0)
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
    xs, ys = shard_batch([x, y], rank, world)
    loss = ((model(xs) - ys) ** 2).mean()           # local mean over this rank's utterances
    loss.backward()                                 # accumulates straight into the flat bucket
    arena.all_reduce_mean(world)
    if rank == 0:
        torch.save({"grad": torch.cat([p.grad.reshape(-1) for p in model.parameters()]),       # the parameters' views of the bucket
                    "flat": torch.cat([p.detach().reshape(-1) for p in model.parameters()]),
                    "in_bucket": all(arena.grad.data_ptr() <= p.grad.data_ptr() < arena.grad.data_ptr() + arena.grad.numel() * 4
                                     for p in model.parameters()),
                    "bucket_sum": float(arena.grad.double().sum())}, out)
    dist.destroy_process_group()


def test_flat_bucket_all_reduce_equals_full_batch_gradient(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, 29531, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
    ((model(x) - y) ** 2).mean().backward()         # equal shards -> mean of local means == global mean
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(got["grad"], ref, atol=1e-6)
    assert torch.equal(got["flat"], torch.cat([p.detach().reshape(-1) for p in model.parameters()]))
    assert got["in_bucket"] and abs(got["bucket_sum"] - float(ref.double().sum())) < 1e-5      # alignment gaps of the bucket stay zero


def _scale_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ml_vae_b200.parallel import global_batch_scale, shard_batch, valid_frame_count
    g = torch.Generator().manual_seed(5)
    T, C = 37, 3
    loss_elem = torch.rand(6, T, C, generator=g) + torch.arange(6.0)[:, None, None]      # unreduced (B, T, C) loss, a different level per utterance
    lens = torch.tensor([1.0, 0.31, 0.5, 0.77, 0.12, 0.93])            # very different numbers of valid frames per rank
    le, ln = shard_batch([loss_elem, lens], rank, world)
    mask = (torch.arange(T)[None, :].float() < ln[:, None] * T).float()[..., None]
    local_mean = (le * mask).sum() / (mask.sum() * C)                  # apply_lens_to_loss(..., 'mean') on this rank's utterances
    f = global_batch_scale(ln, T, world)
    contrib = torch.stack([f * local_mean, local_mean, valid_frame_count(ln, T)])
    dist.all_reduce(contrib, op=dist.ReduceOp.SUM)
    if rank == 0:
        torch.save({"scaled_mean_over_ranks": float(contrib[0] / world), "plain_mean_over_ranks": float(contrib[1] / world),
                    "frames": float(contrib[2])}, out)
    dist.destroy_process_group()


def test_global_batch_scale_recovers_the_global_masked_mean(tmp_path):
    """SURVEY 8e option: averaging (world * cnt_r / cnt_all) * local masked mean over the ranks IS the masked mean of the global batch
    (utils/data_utils.py:67-104 on all utterances at once); the plain average of local means is not when the ranks hold different
    numbers of valid frames."""
    out = str(tmp_path / "s.pt")
    mp.spawn(_scale_worker, args=(2, 29537, out), nprocs=2, join=True)
    got = torch.load(out)
    g = torch.Generator().manual_seed(5)
    T, C = 37, 3
    loss_elem = torch.rand(6, T, C, generator=g) + torch.arange(6.0)[:, None, None]
    lens = torch.tensor([1.0, 0.31, 0.5, 0.77, 0.12, 0.93])
    mask = (torch.arange(T)[None, :].float() < lens[:, None] * T).float()[..., None]
    want = float((loss_elem * mask).sum() / (mask.sum() * C))
    assert got["frames"] == float(mask.sum())
    assert abs(got["scaled_mean_over_ranks"] - want) < 1e-5
    assert abs(got["plain_mean_over_ranks"] - want) > 1e-3


def test_flat_arena_keeps_names_values_and_views():
    from ml_vae_b200.modules import Decoder, VanillaVAE
    from ml_vae_b200.train_step import FlatArena
    torch.manual_seed(3)
    enc, dec = VanillaVAE([12, 8, 8], 4), Decoder(4, 6, 2, 0.0, [12, 8, 8, 12])
    before = {k: v.clone() for k, v in list(enc.state_dict().items()) + list(dec.state_dict().items())}
    arena = FlatArena([enc, dec])
    after = dict(list(enc.state_dict().items()) + list(dec.state_dict().items()))
    assert list(before) == list(after) and all(torch.equal(before[k], after[k]) for k in before)
    total = sum(v.numel() for v in before.values())
    assert total <= arena.flat.numel() < total + 8 * len(before)          # every parameter starts 16-byte aligned (bf16 shadow rows)
    assert all(arena.offset_of(p) % 8 == 0 for p in arena.params)
    assert torch.equal(arena.flat_bf16.float(), arena.flat.bfloat16().float())
    # the heads that run as one stacked GEMM sit back to back: their stacked views are plain slices of the arena
    views = arena.linear_views([enc.mean_fc.weight, enc.log_var_fc.weight], [enc.mean_fc.bias, enc.log_var_fc.bias])
    assert views is None                                                      # latent 4: rows not a multiple of 8 -> generic path
    enc8, dec8 = VanillaVAE([16, 8, 8], 8), Decoder(8, 8, 1, 0.0, [16, 8, 8, 16])
    arena8 = FlatArena([enc8, dec8])
    w16, b32, gw, gb, _ = arena8.linear_views([enc8.mean_fc.weight, enc8.log_var_fc.weight], [enc8.mean_fc.bias, enc8.log_var_fc.bias])
    assert w16.shape == (16, 8) and torch.equal(w16[:8].float(), enc8.mean_fc.weight.detach().bfloat16().float())
    assert torch.equal(w16[8:].float(), enc8.log_var_fc.weight.detach().bfloat16().float()) and torch.equal(b32[8:], enc8.log_var_fc.bias.detach())
    gw.fill_(2.0)
    assert float(enc8.log_var_fc.weight.grad.min()) == 2.0 and float(enc8.mean_fc.weight.grad.min()) == 2.0
    arena.flat.zero_()                               # parameters are views of the arena
    assert all(float(p.abs().sum()) == 0 for p in enc.parameters())
    assert all(p.grad.data_ptr() >= arena.grad.data_ptr() for p in dec.parameters())
