"""Small end-to-end pass over every kernel of the library for compute-sanitizer (memcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_vae_b200.features import Fbank
from ml_vae_b200.modules import Decoder, VanillaVAE
from ml_vae_b200.normalizer import InputNormalization
from ml_vae_b200.train_step import TrainStep
from ml_vae_b200.utils.data_utils import apply_lens_to_loss
dev = torch.device("cuda:0")
torch.manual_seed(0)
for hop, mels, dl in ((10, 80, False), (20, 40, True), (12.5, 40, True)):
    wav = 0.1 * torch.randn(3, 4321 if hop == 12.5 else 4800, device=dev)
    lens = torch.tensor([wav.shape[1], 3000, 801], dtype=torch.int32, device=dev)
    f, r = Fbank(deltas=dl, hop_length=hop, n_mels=mels)(wav, lens, truncate=True)
enc = VanillaVAE([80, 64, 64], 64).to(dev)
dec = Decoder(64, 64, 2, 0.0, [128, 64, 64, 80]).to(dev)
ts = TrainStep(Fbank(deltas=False, hop_length=10, n_mels=80), InputNormalization().to(dev), enc, dec,
               {"kld_weight": 0.001, "batch_size": 5}, compute_dtype=torch.bfloat16)
wav = 0.1 * torch.randn(5, 6400, device=dev)
lens = torch.tensor([6400, 6000, 4000, 3200, 1600], dtype=torch.int32, device=dev)
for _ in range(2):
    loss = ts.step(wav, lens)
x = torch.randn(2, 9, 7, device=dev, requires_grad=True)
apply_lens_to_loss(x, torch.tensor([1.0, 0.5], device=dev), "batch").sum().backward()
torch.cuda.synchronize()
print("sanitize pass done, loss", float(loss))
