"""Probe: which torch (non-library) ops still launch kernels inside one eager bf16 training step, with shapes and call sites."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import ProfilerActivity, profile
from ml_vae_b200.features import Fbank
from ml_vae_b200.modules import Decoder, VanillaVAE
from ml_vae_b200.normalizer import InputNormalization
from ml_vae_b200.train_step import TrainStep
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, n = 64, 80000
ts = TrainStep(Fbank(deltas=True, sample_rate=16000, hop_length=10, n_fft=400, n_mels=80), InputNormalization().to(dev),
               VanillaVAE([240, 64, 64], 64).to(dev), Decoder(64, 512, 2, 0.15, [1024, 64, 64, 240]).to(dev),
               {"kld_weight": 0.001, "batch_size": B}, lr=1e-3)
wav = 0.1 * torch.randn(B, n, device=dev); lens = torch.full((B,), n, dtype=torch.int32, device=dev)
for _ in range(3): ts.step(wav, lens)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    ts.step(wav, lens); torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6):
    t = getattr(e, "self_device_time_total", 0) or 0
    if t > 0 and e.key.startswith("aten::"):
        stack = [s.split("/")[-1] for s in (e.stack or []) if "ml_vae_b200" in s][:2]
        rows.append((t, e.count, e.key, str(e.input_shapes)[:80], stack))
for t, c, k, sh, st in sorted(rows, reverse=True)[:40]:
    print(f"{t:7.1f} us x{c:<2d} {k:26s} {sh:80s} {st}")
