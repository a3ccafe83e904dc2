"""2-GPU check (torchrun --nproc-per-node 2): the overlapped gradient all-reduce (tail of the bucket under the first LSTM
layer's backward) gives bit-identical parameters to the single all-reduce after backward, eager and as a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from ml_vae_b200.features import Fbank
from ml_vae_b200.modules import Decoder, VanillaVAE
from ml_vae_b200.normalizer import InputNormalization
from ml_vae_b200.train_step import TrainStep

dev = torch.device("cuda", local)
B, n = 16, 16000


def run(overlap, graph):
    torch.manual_seed(123456)
    fb = Fbank(deltas=False, hop_length=10, n_mels=80)
    enc = VanillaVAE([80, 64, 64], 64).to(dev)
    dec = Decoder(64, 128, 2, 0.0, [256, 64, 64, 80]).to(dev)
    ts = TrainStep(fb, InputNormalization().to(dev), enc, dec, {"kld_weight": 0.001, "batch_size": B}, world_size=world,
                   overlap_all_reduce=overlap)
    ts.arena.broadcast(0)
    g = torch.Generator().manual_seed(7 + rank)
    wavs = [(0.1 * torch.randn(B, n, generator=g)).to(dev) for _ in range(4)]
    lens = torch.full((B,), n, dtype=torch.int32, device=dev)
    if graph:
        assert ts.capture(wavs[0], lens, warmup=1)
    losses = [float(ts.step(w, lens)) for w in wavs]
    torch.cuda.synchronize()
    used = ts._early_done
    ts._graph = None
    return ts.arena.flat.clone(), losses, used


ref, l0, u0 = run(False, False)
for overlap, graph in ((True, False), (True, True), (False, True)):
    out, l1, used = run(overlap, graph)
    same = torch.equal(ref, out)
    # all ranks must also agree with each other
    other = out.clone()
    dist.broadcast(other, 0)
    agree = torch.equal(other, out)
    if rank == 0:
        print(f"overlap={overlap} graph={graph}: early all-reduce used={used}  params identical to baseline: {same}  ranks agree: {agree}  "
              f"losses {['%.5f' % v for v in l1]}", flush=True)
    assert agree and (same or graph), "overlapped all-reduce changed the result"
if rank == 0:
    print("OVERLAP OK", flush=True)
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
