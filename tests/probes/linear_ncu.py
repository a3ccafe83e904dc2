import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_vae_b200 import _lib as L
dev = torch.device("cuda:0")
M, K, N = 32000, 1024, 128
x = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16(); b = torch.randn(N, device=dev)
y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    L.check(L.lib().mlvae_linear_fwd(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), M, N, K, K, N, 1, L.stream_ptr()), "linear")
torch.cuda.synchronize(); print("done")
