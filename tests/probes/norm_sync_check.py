"""2-rank check (torchrun): InputNormalization(sync_stats=True) keeps identical running statistics on every rank, equal to ONE normaliser
that sees the concatenated batches; also inside a captured CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
from ml_vae_b200.normalizer import InputNormalization

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
B, T, D = 4, 50, 80
sync, whole = InputNormalization(sync_stats=True).to(dev), InputNormalization().to(dev)
ok = True
for it, epoch in enumerate([0, 0, 1, 2, 4]):
    g = torch.Generator().manual_seed(100 + it)
    x_all = torch.randn(world * B, T, D, generator=g) * (1 + it) + it
    lens_all = torch.rand(world * B, generator=g) * 0.5 + 0.5
    x, lens = x_all[rank * B:(rank + 1) * B].to(dev), lens_all[rank * B:(rank + 1) * B].to(dev)
    out = sync(x, lens, epoch=epoch)
    want = whole(x_all.to(dev), lens_all.to(dev), epoch=epoch)[rank * B:(rank + 1) * B]
    err = float((out - want).abs().max() / want.abs().max())
    ok &= err < 1e-5
states = [torch.zeros_like(sync._state) for _ in range(world)]
dist.all_gather(states, sync._state)
same = all(torch.equal(states[0], s) for s in states)
err_m = float((sync.glob_mean - whole.glob_mean).abs().max() / whole.glob_mean.abs().max())
# graph capture of the exchange
xs, ls = x.clone(), lens.clone()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    sync(xs, ls, epoch=0)
    torch.cuda.synchronize()
    with torch.cuda.graph(gr, stream=s):
        o = sync(xs, ls, epoch=0)
    gr.replay(); gr.replay()
torch.cuda.synchronize()
cnt = sync.count
if rank == 0:
    print(f"world {world}: outputs match the global-batch normaliser: {ok}; identical state on all ranks: {same}; running mean rel err {err_m:.1e}; "
          f"count after graph replays {cnt} (whole {whole.count})", flush=True)
sys.stdout.flush()
os._exit(0)      # like bench.py: a captured graph holds NCCL work; tearing the process group down under it can hang
