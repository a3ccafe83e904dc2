"""One mlvae_md_decode launch at a BASELINE-sized batch for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_vae_b200.utils import decode_utils as du
from test_decode_gpu import _random_logs
dev = torch.device("cuda:0")
args = [torch.as_tensor(a).to(dev) for a in _random_logs(1, 64, 500, 42, 60)]
for _ in range(3):
    du.decode_from_logs(*args, device=dev)
torch.cuda.synchronize()
print("done")
