"""Probe: fused Linear backward prologue vs the torch ops it replaces (run under `timeout`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_vae_b200 import dense

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


for M, N in [(32000, 64), (32000, 80), (32000, 128), (32000, 1024), (1 << 20, 64)]:
    dy = torch.randn(M, N, device=dev).bfloat16()
    y = torch.randn(M, N, device=dev).bfloat16()
    for leaky in (False, True):
        ours = t(lambda: dense._bwd_prep(dy, y if leaky else None))

        def ref():
            g = dy * torch.where(y > 0, 1.0, 0.01).to(dy.dtype) if leaky else dy
            return g, g.sum(0, dtype=torch.float32)
        theirs = t(ref)
        byts = M * N * 2 * (3 if leaky else 1)
        print(f"M={M} N={N} leaky={leaky}: ours {ours:7.1f} us ({byts / ours / 1e3:6.0f} GB/s)  torch {theirs:7.1f} us", flush=True)
