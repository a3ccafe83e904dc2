"""Probe: why is dW_ih = dA^T x slow for a narrow input (In = 64)?  Split / debug-mode sweep against cuBLAS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_vae_b200 import _lib as L
from ml_vae_b200.gemm import gemm
dev = torch.device("cuda:0")
Bb, T, H = 64, 500, 512
M = Bb * T
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]
dA = torch.randn(Bb, T, 2, 4 * H, device=dev).bfloat16()
dA2 = dA.view(M, 8 * H)
for In in (64,):
    x = torch.randn(M, In, device=dev).bfloat16()
    dW = [torch.zeros(4 * H, In, device=dev) for _ in range(2)]
    print(f"cuBLAS dA^T x In={In}: {timed(lambda: torch.mm(dA2.t(), x)):7.1f} us", flush=True)
    for mode in (0, 1):
        L.lib().mlvae_gemm_debug_mode(mode)
        for split in (1, 2, 4, 8, 9, 16):
            f = lambda: gemm([dA2[:, :4 * H], dA2[:, 4 * H:]], [x, x], dW, 4 * H, In, M, lda=8 * H, ldb=In, ldd=In, a_mn=True, b_mn=True, out_f32=True,
                             accumulate=True, row_perm_H=H, split_k=split, bn=64)
            print(f"dW_ih In={In} mode={mode} split={split}: {timed(f):7.1f} us", flush=True)
    L.lib().mlvae_gemm_debug_mode(0)
    # one problem over the full 4096 columns (no grouping): same bytes
    dWf = torch.zeros(8 * H, In, device=dev)
    for split in (2, 4, 8):
        f = lambda: gemm(dA2, x, dWf, 8 * H, In, M, lda=8 * H, ldb=In, ldd=In, a_mn=True, b_mn=True, out_f32=True, accumulate=True, split_k=split, bn=64)
        print(f"dW_ih In={In} single problem split={split}: {timed(f):7.1f} us", flush=True)
    # P = x W^T (N = 4096 output, K = 64)
    P = torch.empty(M, 8 * H, dtype=torch.bfloat16, device=dev); w = torch.randn(8 * H, In, device=dev).bfloat16(); bias = torch.randn(8 * H, device=dev)
    print(f"cuBLAS P In={In}: {timed(lambda: torch.addmm(bias.bfloat16(), x, w.t())):7.1f} us", flush=True)
    for mode in (0, 1, 2):
        L.lib().mlvae_gemm_debug_mode(mode)
        for bn in (128, 256):
            print(f"P In={In} mode={mode} bn={bn}: {timed(lambda: gemm(x, w, P, M, 8 * H, In, lda=In, ldb=In, ldd=8 * H, bias=bias, bn=bn)):7.1f} us", flush=True)
    L.lib().mlvae_gemm_debug_mode(0)
    # dx = dA W (N = 64, K = 4096)
    dx = torch.empty(M, In, dtype=torch.bfloat16, device=dev)
    print(f"cuBLAS dx In={In}: {timed(lambda: torch.mm(dA2, w)):7.1f} us", flush=True)
    print(f"dx In={In} bn=64: {timed(lambda: gemm(dA2, w, dx, M, In, 8 * H, lda=8 * H, ldb=In, ldd=In, b_mn=True, bn=64)):7.1f} us", flush=True)
