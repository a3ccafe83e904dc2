import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_vae_b200 import _lib as L
from ml_vae_b200.gemm import gemm
dev = torch.device("cuda:0")
M, H = 32000, 512
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]
L.lib().mlvae_gemm_debug_mode.argtypes = [L.C.c_int]
for In in (1024, 64):
    x = torch.randn(M, In, device=dev).bfloat16(); w = torch.randn(8 * H, In, device=dev).bfloat16(); bias = torch.randn(8 * H, device=dev)
    P = torch.empty(M, 8 * H, dtype=torch.bfloat16, device=dev)
    for mode, name in ((0, "normal"), (2, "no bias"), (1, "no stores")):
        L.lib().mlvae_gemm_debug_mode(mode)
        print(f"In={In} {name:10s} {timed(lambda: gemm(x, w, P, M, 8 * H, In, lda=In, ldb=In, ldd=8 * H, bias=bias, bn=256)):8.1f} us   (bias=None: "
              f"{timed(lambda: gemm(x, w, P, M, 8 * H, In, lda=In, ldb=In, ldd=8 * H, bn=256)):8.1f} us)", flush=True)
L.lib().mlvae_gemm_debug_mode(0)
