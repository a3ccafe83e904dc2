import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_vae_b200 import _lib as L
dev = torch.device("cuda:0")
ok = True
for M, K, N, leaky in [(128, 64, 64, 0), (300, 80, 64, 1), (32000, 80, 64, 1), (32000, 64, 128, 0), (32000, 1024, 128, 1), (32000, 64, 80, 0),
                       (1000, 120, 64, 1), (777, 1024, 64, 1), (32000, 1024, 64, 1), (513, 64, 120, 0), (32000, 128, 1024, 0), (100, 8, 16, 0)]:
    if N > 256: continue
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).bfloat16().to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(dev)
    b = torch.randn(N, generator=g).to(dev)
    y = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    L.check(L.lib().mlvae_linear_fwd(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), M, N, K, K, N, leaky, L.stream_ptr()), "linear")
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b
    if leaky: ref = torch.nn.functional.leaky_relu(ref, 0.01)
    err = float((y.float() - ref).abs().max() / ref.abs().max())
    t, tt = [], []
    bb = b.bfloat16()
    lib, args, st = L.lib(), (L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), M, N, K, K, N, leaky), L.stream_ptr()
    for _ in range(3):
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(20): lib.mlvae_linear_fwd(*args, st)
        b_.record(); torch.cuda.synchronize(); t.append(a_.elapsed_time(b_) / 20)
        a_.record()
        for _ in range(20): z = torch.nn.functional.linear(x, w, bb)
        b_.record(); torch.cuda.synchronize(); tt.append(a_.elapsed_time(b_) / 20)
    gbs = (M * K * 2 + M * N * 2) / (min(t) * 1e-3) / 1e9
    print(f"M={M} K={K} N={N} leaky={leaky}: rel err {err:.2e} {'OK' if err < 8e-3 else 'MISMATCH'}  {min(t)*1e3:.1f} us ({gbs:.0f} GB/s) vs cuBLAS {min(tt)*1e3:.1f} us", flush=True)
    ok &= err < 8e-3
print("ALL OK" if ok else "FAILED")
