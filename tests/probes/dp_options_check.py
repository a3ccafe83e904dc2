"""2-rank check (torchrun): TrainStep with the two optional exchanges of SURVEY 8e switched on -- normaliser statistics
(InputNormalization(sync_stats=True)) and the exact global-batch masked mean (global_batch_mean=True) -- eager steps, then captured
in the CUDA graph and replayed; ragged lengths that differ between the ranks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
from ml_vae_b200.features import Fbank
from ml_vae_b200.modules import Decoder, VanillaVAE
from ml_vae_b200.normalizer import InputNormalization
from ml_vae_b200.train_step import TrainStep

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(1)
B, n = 16, 32000
enc = VanillaVAE([80, 64, 64], 64).to(dev)
dec = Decoder(64, 128, 2, 0.15, [256, 64, 64, 80]).to(dev)
norm = InputNormalization(sync_stats=True).to(dev)
ts = TrainStep(Fbank(deltas=False, hop_length=10, n_mels=80), norm, enc, dec, {"kld_weight": 0.001, "batch_size": B},
               world_size=world, global_batch_mean=True)
g = torch.Generator().manual_seed(7 + rank)
wav = (0.1 * torch.randn(B, n, generator=g)).to(dev)
lens = torch.randint(n // (2 + 2 * rank), n + 1, (B,), generator=g).to(torch.int32).to(dev)      # rank 1 holds shorter utterances
lens[0] = n
losses = [float(ts.step(wav, lens)) for _ in range(3)]
graphed = ts.capture(wav, lens, warmup=1)
losses += [float(ts.step(wav, lens)) for _ in range(3)]
torch.cuda.synchronize()
st = [torch.zeros_like(norm._state) for _ in range(world)]
dist.all_gather(st, norm._state)
flat = [torch.zeros_like(ts.arena.flat) for _ in range(world)]
dist.all_gather(flat, ts.arena.flat)
if rank == 0:
    print(f"world {world} dp_peer {ts.dp_peer} graph {graphed}: losses {['%.4f' % v for v in losses]}; finite {all(v == v for v in losses)}; "
          f"normaliser state identical on all ranks {all(torch.equal(st[0], s) for s in st)}; parameters identical {all(torch.equal(flat[0], f) for f in flat)}", flush=True)
sys.stdout.flush()
os._exit(0)      # like bench.py: a captured graph holds NCCL work; tearing the process group down under it can hang
