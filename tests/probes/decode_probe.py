"""Probe: mlvae_md_decode (one CTA per utterance) vs the CPU port of the reference's DP (oracle/decode_ref.py, numpy-vectorised over
the phoneme axis -- the reference itself is a python triple loop under joblib, slower still) at BASELINE-sized batches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from ml_vae_b200.utils import decode_utils as du
from oracle import decode_ref
from test_decode_gpu import _random_logs

dev = torch.device("cuda:0")
for B, T, N, Lmax in [(64, 500, 42, 60), (16, 2000, 42, 100), (64, 500, 42, 200)]:
    args = _random_logs(1, B, T, N, Lmax)
    targs = [torch.as_tensor(a).to(dev) for a in args]
    for _ in range(3):
        out = du.decode_from_logs(*targs, device=dev)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        out = du.decode_from_logs(*targs, device=dev)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    n_cpu = min(B, 4)
    t0 = time.perf_counter()
    ob, of, op = decode_ref.decode_batch(*[x[:n_cpu] if i not in (3,) else x for i, x in enumerate(args)])
    cpu_ms = (time.perf_counter() - t0) * 1e3 / n_cpu * B
    ok = all(np.array_equal(out[0][i, :args[5][i]].cpu().numpy(), ob[i]) for i in range(n_cpu))
    print(f"B={B} T={T} N={N} Lmax={Lmax}: GPU {ms:.3f} ms per batch ({B / ms * 1e3:.0f} utt/s), CPU port {cpu_ms:.0f} ms per batch "
          f"(1 core, extrapolated from {n_cpu} utterances) -> x{cpu_ms / ms:.0f}; first {n_cpu} utterances identical: {ok}", flush=True)
