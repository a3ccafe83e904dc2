"""Probe: per-STEP phase durations of one gate warp of the persistent LSTM kernels (instrumented instantiation): is the exchange wait a
narrow distribution or are there late steps, and do they come with a period?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from ml_vae_b200 import _lib as L
from ml_vae_b200.lstm import bilstm_layer

dev = torch.device("cuda:0")
B, T, In, H = int(os.environ.get("LSTM_B", 64)), 500, 64, 512
torch.manual_seed(0)
ref = torch.nn.LSTM(In, H, 1, bidirectional=True, batch_first=True).to(dev)
names = [f"{k}_l0{s}" for s in ("", "_reverse") for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
ps = [getattr(ref, n).detach().clone().requires_grad_(True) for n in names]
x = torch.randn(B, T, In, device=dev).bfloat16().requires_grad_(True)
gy = torch.randn(B, T, 2 * H, device=dev).bfloat16()
lib = L.lib()
lib.mlvae_debug_set_trace_buffer.argtypes = [L.C.c_void_p]
if os.environ.get("BWD_DELAY"): lib.mlvae_debug_set_option(4, int(os.environ["BWD_DELAY"]))
for _ in range(2):
    bilstm_layer(x, *ps, training=True).backward(gy)
torch.cuda.synchronize()
prof = torch.zeros(128, dtype=torch.int64, device=dev)
for which in ("fwd", "bwd"):
    trace = torch.zeros(4 * T, dtype=torch.int32, device=dev)
    y = None
    if which == "bwd":
        y = bilstm_layer(x, *ps, training=True)
    L.check(lib.mlvae_debug_set_profile_buffer(L.ptr(prof)), "prof")
    L.check(lib.mlvae_debug_set_trace_buffer(L.ptr(trace)), "trace")
    if which == "fwd":
        y = bilstm_layer(x, *ps, training=True)
        torch.cuda.synchronize()
        L.check(lib.mlvae_debug_set_trace_buffer(None), "trace")
        L.check(lib.mlvae_debug_set_profile_buffer(None), "prof")
    else:
        y.backward(gy)
        torch.cuda.synchronize()
        L.check(lib.mlvae_debug_set_trace_buffer(None), "trace")
        L.check(lib.mlvae_debug_set_profile_buffer(None), "prof")
    tr = trace.cpu().numpy().reshape(T, 4)[2:T - 2]
    tot = tr.sum(1)
    print(f"{which}: per-step cycles, {len(tr)} steps: total mean {tot.mean():.0f}  p10 {np.percentile(tot, 10):.0f}  p50 {np.percentile(tot, 50):.0f}  "
          f"p90 {np.percentile(tot, 90):.0f}  p99 {np.percentile(tot, 99):.0f}  max {tot.max()}")
    for k, name in enumerate(["exchange wait", "phase 1", "phase 2", "phase 3"]):
        v = tr[:, k]
        print(f"   {name:14s} mean {v.mean():6.0f}  p10 {np.percentile(v, 10):6.0f}  p50 {np.percentile(v, 50):6.0f}  p90 {np.percentile(v, 90):6.0f}  p99 {np.percentile(v, 99):6.0f}  max {v.max():6d}")
    w = tr[:, 0]
    hist, edges = np.histogram(w, bins=[0, 400, 600, 700, 800, 900, 1000, 1100, 1200, 1400, 1600, 2000, 3000, 100000])
    print("   exchange-wait histogram:", " ".join(f"<{int(e)}:{h}" for h, e in zip(hist, edges[1:])))
    late = np.nonzero(w > np.percentile(w, 50) + 300)[0]
    print(f"   late steps (> p50 + 300): {len(late)}; gaps between them: {np.diff(late)[:30].tolist()}")
    print("   first 40 waits:", w[:40].tolist())
    par = (np.arange(2, T - 2) & 1)
    for q in (0, 1):
        m = tr[par == q].mean(0)
        print(f"   steps with (step & 1) == {q}: wait {m[0]:.0f}  p1 {m[1]:.0f}  p2 {m[2]:.0f}  p3 {m[3]:.0f}  total {m.sum():.0f}")
    print("   steps 100..123 [wait, p1, p2, p3]:", tr[98:122].tolist())
