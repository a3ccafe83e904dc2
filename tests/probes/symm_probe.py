"""Probe (torchrun --nproc-per-node N): what does torch's symmetric memory give on this box? (peer pointers, multicast)"""
import os, sys, time
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as sm
try:
    t = sm.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(rank + 1)
    h = sm.rendezvous(t, dist.group.WORLD)
    print(f"[{rank}] backend {sm.get_backend(dev)} buffer_ptrs {[hex(p) for p in h.buffer_ptrs]} multicast {hex(h.multicast_ptr)} "
          f"signal_pad {h.signal_pad_size} ptrs {[hex(p) for p in h.signal_pad_ptrs]} x", flush=True)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.float32)
    print(f"[{rank}] peer value {peer[:2].tolist()}", flush=True)
    h.barrier()
except Exception as e:
    print(f"[{rank}] symmetric memory failed: {type(e).__name__}: {e}", flush=True)
# plain CUDA IPC through torch storage sharing
try:
    x = torch.full((1 << 20,), float(rank + 1), device=dev)
    info = x.untyped_storage()._share_cuda_()
    print(f"[{rank}] _share_cuda_ -> device {info[0]} handle {len(info[1])} B size {info[2]} offset {info[3]}", flush=True)
except Exception as e:
    print(f"[{rank}] _share_cuda_ failed: {e}", flush=True)
print(f"[{rank}] can_device_access_peer: {[torch.cuda.can_device_access_peer(dev.index, j) for j in range(world) if j != dev.index]}", flush=True)
dist.barrier(); dist.destroy_process_group()
