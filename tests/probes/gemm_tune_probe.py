"""Probe: tile / split-K choices for the LSTM weight-gradient GEMMs (B=64, T=500, H=512)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
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
dA = torch.randn(Bb, T, 2, 4 * H, device=dev).bfloat16(); y = torch.randn(Bb, T, 2 * H, device=dev).bfloat16()
dA2, y2 = dA.view(M, 8 * H), y.view(M, 2 * H)
out = [torch.zeros(4 * H, H, device=dev) for _ in range(2)]
for bn, split in ((128, 1), (256, 1), (256, 2), (256, 3), (128, 2)):
    f = lambda: gemm([dA2[1:, :4 * H], dA2[:, 4 * H:]], [y2[:, :H], y2[1:, H:]], out, 4 * H, H, T - 1, lda=8 * H, ldb=2 * H, ldd=H, a_mn=True, b_mn=True,
                     kbatches=Bb, a_batch_stride=T * 8 * H, b_batch_stride=T * 2 * H, out_f32=True, accumulate=True, row_perm_H=H, split_k=split, bn=bn)
    print(f"dW_hh bn={bn} split={split}: {timed(f):7.1f} us", flush=True)
for In in (64, 256):
    x = torch.randn(M, In, device=dev).bfloat16()
    dW = [torch.zeros(4 * H, In, device=dev) for _ in range(2)]
    for bn, split in ((64, 1), (64, 2), (64, 4), (64, 8), (128, 4), (256, 4)):
        if bn > max(64, In): continue
        f = lambda: gemm([dA2[:, :4 * H], dA2[:, 4 * H:]], [x, x], dW, 4 * H, In, M, lda=8 * H, ldb=In, ldd=In, a_mn=True, b_mn=True, out_f32=True, accumulate=True,
                         row_perm_H=H, split_k=split, bn=bn)
        print(f"dW_ih In={In} bn={bn} split={split}: {timed(f):7.1f} us", flush=True)
    P = torch.empty(M, 8 * H, dtype=torch.bfloat16, device=dev); w = torch.randn(8 * H, In, device=dev).bfloat16(); bias = torch.randn(8 * H, device=dev)
    for bn in (128, 256):
        print(f"P In={In} bn={bn}: {timed(lambda: gemm(x, w, P, M, 8 * H, In, lda=In, ldb=In, ldd=8 * H, bias=bias, bn=bn)):7.1f} us", flush=True)
