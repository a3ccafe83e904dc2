"""Probe (torchrun --nproc-per-node N): time of the optimiser exchange alone on the benchmark arena (8.66 M parameters):
NCCL all-reduce + mlvae_adam_clip_step  vs  mlvae_dp_adam_step over peer memory (P2P and multicast)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from ml_vae_b200 import _lib as L
from ml_vae_b200.peer import PeerArenaMemory
n = int(os.environ.get("DP_N", 8_660_000)) // 8 * 8
lib = L.lib()
loss = torch.ones(1, device=dev)
m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)

def timed(fn, iters=50):
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

# NCCL path
p, g, p16 = torch.randn(n, device=dev), torch.randn(n, device=dev), torch.zeros(n, device=dev, dtype=torch.bfloat16)
state = torch.zeros(lib.mlvae_adam_state_bytes() // 4, device=dev)
def nccl_step():
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
    L.check(lib.mlvae_adam_clip_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), L.ptr(p16), n, 1.0 / world, 1e-3, 0.9, 0.999, 1e-8, 5.0, L.ptr(state), L.ptr(loss), L.stream_ptr()), "adam", kernels=2)
def nccl_only():
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
def adam_only():
    L.check(lib.mlvae_adam_clip_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), L.ptr(p16), n, 1.0 / world, 1e-3, 0.9, 0.999, 1e-8, 5.0, L.ptr(state), L.ptr(loss), L.stream_ptr()), "adam", kernels=2)
t_ar, t_ad, t_both = timed(nccl_only), timed(adam_only), timed(nccl_step)
if rank == 0:
    print(f"world {world}, arena {n} f32 ({n * 4 / 1e6:.1f} MB): NCCL all-reduce {t_ar:.1f} us, full Adam {t_ad:.1f} us, both {t_both:.1f} us", flush=True)
for mc in (False, True):
    mem = PeerArenaMemory(n, dev, multicast=mc)
    mem.grad.normal_(); mem.flat.normal_()
    a = L.DpAdamArgs(); mem.fill_args(a)
    a.exp_avg, a.exp_avg_sq = m.data_ptr(), v.data_ptr()
    a.lr, a.beta1, a.beta2, a.eps, a.max_grad_norm = 1e-3, 0.9, 0.999, 1e-8, 5.0
    a.loss = loss.data_ptr()
    def peer_step():
        L.check(lib.mlvae_dp_adam_step(a, L.stream_ptr()), "dp", kernels=2)
    t = timed(peer_step)
    st = mem.read_state()
    if rank == 0:
        print(f"peer-memory step ({'multicast' if mem.multicast_base else 'P2P'}): {t:.1f} us  state {st}", flush=True)
    # the two kernels separately (events around each launch are not possible through one C call: use the profiler)
    from torch.profiler import ProfilerActivity, profile
    dist.barrier(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10): peer_step()
        torch.cuda.synchronize()
    if rank == 0:
        for e in prof.key_averages():
            if "dp_" in e.key:
                print(f"    {e.key[:60]:60s} {e.device_time_total / e.count:8.1f} us x{e.count}", flush=True)
    del mem
dist.barrier(); torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
