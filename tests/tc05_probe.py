"""Stand-alone probe of the tcgen05 tile (run under `timeout`): prints max error vs torch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_vae_b200 import _lib as L

dev = torch.device("cuda:0")
ok = True
for N, K, ts in [(64, 64, 0), (16, 16, 0), (128, 64, 0), (256, 128, 0), (80, 64, 0), (64, 256, 0), (144, 320, 0), (256, 256, 0),
                 (16, 16, 1), (16, 64, 1), (32, 512, 1), (64, 256, 1), (16, 512, 1)]:
    g = torch.Generator().manual_seed(N * 1000 + K)
    a = torch.randn(128, K, generator=g).bfloat16().to(dev)
    b = torch.randn(N, K, generator=g).bfloat16().to(dev)
    d = torch.full((128, N), float("nan"), device=dev)
    L.check(L.lib().mlvae_tc05_selftest(L.ptr(a), L.ptr(b), L.ptr(d), N, K, ts, L.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    err = float((d - ref).abs().max() / ref.abs().max())
    print(f"N={N:4d} K={K:4d} A-in-TMEM={ts} rel err {err:.3e}", "OK" if err < 1e-5 else "MISMATCH", flush=True)
    ok &= err < 1e-5
print("ALL OK" if ok else "FAILED")
