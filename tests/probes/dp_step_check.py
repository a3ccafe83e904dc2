"""Probe (torchrun --nproc-per-node N): the peer-memory optimiser step against the NCCL all-reduce + full Adam path.
Same seeds, same batches: after K steps the parameters of the two TrainSteps must agree to float32 round-off, every rank
must hold bit-identical parameters, and the step time of both modes is printed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from ml_vae_b200.features import Fbank
from ml_vae_b200.modules import Decoder, VanillaVAE
from ml_vae_b200.normalizer import InputNormalization
from ml_vae_b200.train_step import TrainStep

B, n, K = 64, 80000, int(os.environ.get("DP_STEPS", 6))

def build(mode):
    if mode == "local":          # no exchange at all: what one GPU of THIS box does alone (same-box baseline for the scaling loss)
        torch.manual_seed(123456)
        fb = Fbank(deltas=True, sample_rate=16000, hop_length=10, n_fft=400, n_mels=80)
        enc = VanillaVAE([240, 64, 64], 64).to(dev)
        dec = Decoder(64, 512, 2, 0.15, [1024, 64, 64, 240]).to(dev)
        return TrainStep(fb, InputNormalization().to(dev), enc, dec, {"kld_weight": 0.001, "batch_size": B}, lr=1e-3, world_size=1)
    torch.manual_seed(123456)
    fb = Fbank(deltas=True, sample_rate=16000, hop_length=10, n_fft=400, n_mels=80)
    enc = VanillaVAE([240, 64, 64], 64).to(dev)
    dec = Decoder(64, 512, 2, 0.15, [1024, 64, 64, 240]).to(dev)
    return TrainStep(fb, InputNormalization().to(dev), enc, dec, {"kld_weight": 0.001, "batch_size": B}, lr=1e-3, world_size=world, dp_mode=mode)

g = torch.Generator().manual_seed(1000 + rank)
wavs = [(0.1 * torch.randn(B, n, generator=g)).to(dev) for _ in range(4)]
lens = torch.full((B,), n, dtype=torch.int32, device=dev)
res = {}
for mode in ("nccl", "peer"):
    ts = build(mode)
    losses = [float(ts.step(wavs[i % 4], lens)) for i in range(K)]
    torch.cuda.synchronize()
    res[mode] = (ts.arena.flat.clone(), losses, ts)
    if mode == "peer":
        st = ts.arena.peer.read_state()
        print(f"[{rank}] peer state {st} multicast {bool(ts.arena.peer.multicast_base)}", flush=True)
    # bit-identical across ranks?
    ref = ts.arena.flat.clone(); dist.broadcast(ref, 0)
    same = torch.equal(ref, ts.arena.flat)
    flag = torch.tensor([int(same)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{mode}: losses {['%.6f' % l for l in losses]}  parameters identical on all ranks: {bool(flag.item())}", flush=True)
pa, pb = res["nccl"][0], res["peer"][0]
d = float((pa - pb).abs().max()); rel = d / float(pa.abs().max())
if rank == 0:
    print(f"max |param_nccl - param_peer| after {K} steps = {d:.3e} (rel {rel:.2e}); loss diff {max(abs(a - b) for a, b in zip(res['nccl'][1], res['peer'][1])):.2e}", flush=True)
# timing, CUDA-graph replays, both modes on the same box
res["local"] = (None, None, build("local"))
for mode in ("local", "nccl", "peer", "local", "nccl", "peer"):
    ts = res[mode][2]
    if ts._graph is None:
        ts.capture(wavs[0], lens, warmup=2)
    for i in range(5):
        ts.step(wavs[i % 4], lens)
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(30):
        ts.step(wavs[i % 4], lens)
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 30], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{mode}: {float(t):.3f} ms/step (graph={ts._graph is not None}, max over ranks, no L2 flush)", flush=True)
if res["peer"][2].dp_peer:
    print(f"[{rank}] final peer state {res['peer'][2].arena.peer.read_state()}", flush=True)
for m_ in res.values():
    m_[2]._graph = None
dist.barrier(); torch.cuda.synchronize()
sys.stdout.flush(); os._exit(0)
