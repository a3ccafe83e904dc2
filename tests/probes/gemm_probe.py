"""Probe: csrc/gemm.cu vs torch (cuBLAS) on the GEMM shapes of one training step (B=64, T=500, H=512).  Run under timeout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_vae_b200.gemm import gemm

dev = torch.device("cuda:0")
torch.manual_seed(0)
Bb, T, H = 64, 500, 512
M = Bb * T
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def report(name, flops, ours, lib, err):
    print(f"{name:44s} ours {ours:8.1f} us ({flops / ours / 1e6:7.1f} TF/s)   cuBLAS {lib:8.1f} us ({flops / lib / 1e6:7.1f} TF/s)   ratio {ours / lib:5.2f}   err {err:.1e}", flush=True)


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max())


for In in (1024, 64):
    x = torch.randn(M, In, device=dev).bfloat16()
    w = (torch.randn(8 * H, In, device=dev) / In ** 0.5).bfloat16()
    bias = torch.randn(8 * H, device=dev)
    P = torch.empty(M, 8 * H, dtype=torch.bfloat16, device=dev)
    bias16 = bias.bfloat16()
    for bn in (256, 128):
        ours = timed(lambda: gemm(x, w, P, M, 8 * H, In, lda=In, ldb=In, ldd=8 * H, bias=bias, bn=bn))
        ref = torch.addmm(bias16, x, w.t())
        lib = timed(lambda: torch.addmm(bias16, x, w.t()))
        report(f"P = x W_ih^T + b   In={In} bn={bn}", 2.0 * M * 8 * H * In, ours, lib, rel(P, ref))
    dA = torch.randn(M, 8 * H, device=dev).bfloat16()
    dx = torch.empty(M, In, dtype=torch.bfloat16, device=dev)
    for bn in ((256, 128) if In > 64 else (64,)):
        ours = timed(lambda: gemm(dA, w, dx, M, In, 8 * H, lda=8 * H, ldb=In, ldd=In, b_mn=True, bn=bn))
        lib = timed(lambda: dA @ w)
        report(f"dx = dA W_ih       In={In} bn={bn}", 2.0 * M * 8 * H * In, ours, lib, rel(dx, dA @ w))
    dW = [torch.zeros(4 * H, In, device=dev) for _ in range(2)]
    for bn in ((256, 128) if In > 64 else (64,)):
        def f():
            gemm([dA[:, :4 * H], dA[:, 4 * H:]], [x, x], dW, 4 * H, In, M, lda=8 * H, ldb=In, ldd=In, a_mn=True, b_mn=True, out_f32=True, bn=bn)
        ours = timed(f)
        lib = timed(lambda: torch.mm(dA.t(), x, out_dtype=torch.float32))
        ref = torch.mm(dA.t(), x, out_dtype=torch.float32)
        report(f"dW_ih = dA^T x (2 dirs grouped) In={In} bn={bn}", 2.0 * M * 8 * H * In, ours, lib, rel(torch.cat(dW), ref))

dA = torch.randn(Bb, T, 2, 4 * H, device=dev).bfloat16()
y = torch.randn(Bb, T, 2 * H, device=dev).bfloat16()
dA2, y2 = dA.view(M, 8 * H), y.view(M, 2 * H)
out = [torch.zeros(4 * H, H, device=dev) for _ in range(2)]
for bn in (128, 256):
    def f():
        gemm([dA2[1:, :4 * H], dA2[:, 4 * H:]], [y2[:, :H], y2[1:, H:]], out, 4 * H, H, T - 1, lda=8 * H, ldb=2 * H, ldd=H, a_mn=True, b_mn=True,
             kbatches=Bb, a_batch_stride=T * 8 * H, b_batch_stride=T * 2 * H, out_f32=True, bn=bn)
    ours = timed(f)
    def g():
        g0 = torch.mm(dA2[1:, :4 * H].t(), y2[:-1, :H], out_dtype=torch.float32)
        g1 = torch.mm(dA2[:-1, 4 * H:].t(), y2[1:, H:], out_dtype=torch.float32)
        g0 -= torch.mm(dA[1:, 0, 0].t(), y[:-1, T - 1, :H], out_dtype=torch.float32)
        g1 -= torch.mm(dA[:-1, T - 1, 1].t(), y[1:, 0, H:], out_dtype=torch.float32)
        return g0, g1
    lib = timed(g)
    g0, g1 = g()
    report(f"dW_hh (2 dirs, shifted, batched) bn={bn}", 2.0 * 2 * 4 * H * H * Bb * (T - 1), ours, lib, max(rel(out[0], g0), rel(out[1], g1)))

# dense layers: weight gradients (split-K) and the wide forward Linear
for (No, Ki) in ((128, 1024), (64, 64), (80, 64), (128, 64)):
    gy = torch.randn(M, No, device=dev).bfloat16()
    x = torch.randn(M, Ki, device=dev).bfloat16()
    dW = torch.zeros(No, Ki, device=dev)
    for split in (8, 16, 32, 64):
        ours = timed(lambda: gemm(gy, x, dW, No, Ki, M, lda=No, ldb=Ki, ldd=Ki, a_mn=True, b_mn=True, out_f32=True, split_k=split))
        lib = timed(lambda: torch.mm(gy.t(), x, out_dtype=torch.float32))
        report(f"dense dW = g^T x  {No}x{Ki} split={split}", 2.0 * M * No * Ki, ours, lib, rel(dW, torch.mm(gy.t(), x, out_dtype=torch.float32)))
x = torch.randn(M, 1024, device=dev).bfloat16()
w = (torch.randn(128, 1024, device=dev) / 32).bfloat16()
b = torch.randn(128, device=dev)
yy = torch.empty(M, 128, dtype=torch.bfloat16, device=dev)
ours = timed(lambda: gemm(x, w, yy, M, 128, 1024, lda=1024, ldb=1024, ldd=128, bias=b, leaky=True))
b16 = b.bfloat16()
lib = timed(lambda: torch.nn.functional.leaky_relu(torch.addmm(b16, x, w.t()), 0.01))
report("Linear fwd 32000x1024 -> 128 (+bias, leaky)", 2.0 * M * 128 * 1024, ours, lib, rel(yy, torch.nn.functional.leaky_relu(torch.addmm(b16, x, w.t()), 0.01)))
