#!/usr/bin/env python
"""Per-kernel roofline microbenchmarks (BASELINE.json configs[4] sweep + the front-end).

    python bench_kernels.py [--quick] [--out profiles/rNN_kernels.json]

Every kernel is timed alone with CUDA events on the launching stream, after 3 warm-ups, on
inputs larger than L2 (126 MB) or with a 256 MiB flush write between launches; the achieved
figure is ALGORITHMIC bytes / time, the denominator is the measured HBM copy
bandwidth in MEASURED_PEAKS.json.  Sweep cap (stated as SURVEY asks): shapes with more than
2^31 elements per tensor are skipped so that five live tensors fit comfortably in HBM.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def time_ms(fn, flush, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--embedded", action="store_true", help="the subset bench.py embeds in its JSON line (configs[4] corners, a few seconds)")
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from ml_vae_b200 import _lib as L
    from ml_vae_b200 import ops
    from ml_vae_b200.build import build
    from ml_vae_b200.features import Fbank
    build()
    dev = torch.device("cuda:0")
    peak = peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []

    def rec(name, shape, dtype, bytes_, ms_med, ms_min):
        gbs = bytes_ / (ms_med * 1e-3) / 1e9
        rows.append({"kernel": name, "shape": shape, "dtype": dtype, "algorithmic_bytes": bytes_,
                     "ms_median": round(ms_med, 5), "ms_min": round(ms_min, 5), "achieved_gbs": round(gbs, 1),
                     "frac_of_measured_peak": round(gbs / peak, 4)})
        print(f"{name:28s} {str(shape):22s} {dtype:5s} {ms_med:9.4f} ms  {gbs:8.1f} GB/s  {gbs / peak:6.1%}", flush=True)

    # BASELINE configs[4]: latent dim 32-1024 x 1M-64M frames, capped at 2^31 elements per tensor (five live tensors)
    lat = [(32, 1 << 20), (32, 16 << 20), (32, 64 << 20), (64, 1 << 20), (64, 4 << 20), (64, 16 << 20), (128, 1 << 20), (128, 4 << 20),
           (128, 16 << 20), (256, 1 << 20), (256, 4 << 20), (512, 1 << 20), (512, 4 << 20), (1024, 1 << 20)]
    if args.quick:
        lat = [(64, 1 << 20), (64, 4 << 20), (256, 1 << 20)]
    if args.embedded:
        lat = [(32, 64 << 20), (64, 1 << 20), (64, 16 << 20), (128, 4 << 20), (512, 4 << 20), (1024, 1 << 20)]
    if not args.only or "latent" in args.only:
        for dt, s, name in ((torch.bfloat16, 2, "bf16"), (torch.float32, 4, "f32")):
            for Ld, M in lat:
                if M * Ld > (1 << 31) or (s == 4 and M * Ld > (1 << 30)):
                    continue
                B, T = 64, M // 64
                mu = torch.randn(B, T, Ld, device=dev, dtype=dt)
                lv = torch.randn(B, T, Ld, device=dev, dtype=dt).clamp_(-6, 3)
                gz = torch.randn(B, T, Ld, device=dev, dtype=dt)
                lens = torch.linspace(0.5, 1.0, B, device=dev)
                z = torch.empty_like(mu); gmu = torch.empty_like(mu); glv = torch.empty_like(mu)
                out = torch.empty(3, device=dev)
                one = torch.ones((), device=dev)
                sc = L.reduce_scratch(dev)
                code = L.dtype_code(mu)
                lib = L.lib()
                st = L.stream_ptr()

                def fwd():
                    L.check(lib.mlvae_reparam_kl_fwd(L.ptr(mu), L.ptr(lv), None, 1, 0, None, L.ptr(lens), B, T, Ld, code,
                                                     L.ptr(z), None, L.ptr(out), L.ptr(sc), st))

                def bwd():
                    L.check(lib.mlvae_reparam_kl_bwd(L.ptr(mu), L.ptr(lv), None, 1, 0, None, L.ptr(gz), None, L.ptr(one),
                                                     L.ptr(lens), B, T, Ld, code, L.ptr(gmu), L.ptr(glv), st))

                big = M * Ld * s * 3 > (200 << 20)
                m, mn = time_ms(fwd, None if big else flush)
                rec("reparam_kl_fwd(philox)", (M, Ld), name, 3 * M * Ld * s, m, mn)
                m, mn = time_ms(bwd, None if big else flush)
                rec("reparam_kl_bwd(philox)", (M, Ld), name, 5 * M * Ld * s, m, mn)
                # reconstruction loss on the same buffers (mean=mu, logvar=lv, target=gz)
                def rfwd():
                    L.check(lib.mlvae_recon_fwd(L.ptr(mu), L.ptr(lv), L.ptr(gz), L.ptr(lens), B, T, Ld, code, 0, None,
                                                L.ptr(out), L.ptr(sc), st))

                def rbwd():
                    L.check(lib.mlvae_recon_bwd(L.ptr(mu), L.ptr(lv), L.ptr(gz), None, L.ptr(one), L.ptr(lens), B, T, Ld,
                                                code, 0, L.ptr(gmu), L.ptr(glv), None, st))
                m, mn = time_ms(rfwd, None if big else flush)
                rec("recon_nll_fwd", (M, Ld), name, 3 * M * Ld * s, m, mn)
                m, mn = time_ms(rbwd, None if big else flush)
                rec("recon_nll_bwd", (M, Ld), name, 5 * M * Ld * s, m, mn)
                del mu, lv, gz, z, gmu, glv
                torch.cuda.empty_cache()

    if args.embedded:
        if args.out:
            os.makedirs(os.path.dirname(os.path.join(ROOT, args.out)) or ".", exist_ok=True)
            json.dump({"peak_gbs": peak, "rows": rows}, open(os.path.join(ROOT, args.out), "w"), indent=1)
        print("EMBEDDED_JSON " + json.dumps({"peak_gbs": peak, "cap": "shapes with more than 2^31 elements per tensor are skipped", "rows": rows}), flush=True)
        return
    if not args.only or "fbank" in args.only:
        for (B, secs, hop_ms, mels, dl, od) in [(64, 5, 10, 80, False, torch.float32), (64, 5, 10, 80, True, torch.bfloat16),
                                                (16, 20, 10, 80, False, torch.float32), (512, 5, 10, 80, False, torch.float32),
                                                (512, 5, 20, 40, True, torch.float32)]:
            n = secs * 16000
            wav = 0.1 * torch.randn(B, n, device=dev)
            fb = Fbank(deltas=dl, hop_length=hop_ms, n_mels=mels)
            lens = torch.full((B,), n, dtype=torch.int32, device=dev)
            hop = 16 * hop_ms
            T = n // hop
            D = mels * (3 if dl else 1)
            s = 2 if od == torch.bfloat16 else 4
            f = lambda: fb(wav, lens, truncate=True, out_dtype=od)
            m, mn = time_ms(f, flush)
            rec("fbank(logmel+finish)", (B, n, f"hop{hop}", f"D{D}"), "bf16" if s == 2 else "f32",
                B * n * 4 + B * T * D * s, m, mn)
            del wav

    if not args.only or "hvae" in args.only:
        # GMM-VAE / H-VAE family: KL against a learned prior + reparameterise, mixture selection
        for dt, sz, name in ((torch.bfloat16, 2, "bf16"), (torch.float32, 4, "f32")):
            for M, NL in ([(1 << 20, 3 * 64)] if args.quick else [(1 << 20, 3 * 64), (4 << 20, 3 * 64), (1 << 20, 8 * 128)]):
                ts = [torch.randn(M, NL, device=dev, dtype=dt).clamp_(-3, 3) for _ in range(6)]
                mu, lv, pmu, plv, gz, gk = ts
                z, kl = torch.empty_like(mu), torch.empty_like(mu)
                outs = [torch.empty_like(mu) for _ in range(4)]
                code, lib, st = L.dtype_code(mu), L.lib(), L.stream_ptr()

                def gfwd():
                    L.check(lib.mlvae_gmm_reparam_kl_fwd(L.ptr(mu), L.ptr(lv), L.ptr(pmu), L.ptr(plv), None, 1, 0, None, mu.numel(),
                                                         code, L.ptr(z), L.ptr(kl), st))

                def gbwd():
                    L.check(lib.mlvae_gmm_reparam_kl_bwd(L.ptr(mu), L.ptr(lv), L.ptr(pmu), L.ptr(plv), None, 1, 0, None, L.ptr(gz),
                                                         L.ptr(gk), mu.numel(), code, *[L.ptr(o) for o in outs], st))
                big = M * NL * sz * 6 > (200 << 20)
                m, mn = time_ms(gfwd, None if big else flush)
                rec("gmm_reparam_kl_fwd(philox)", (M, NL), name, 6 * M * NL * sz, m, mn)
                m, mn = time_ms(gbwd, None if big else flush)
                rec("gmm_reparam_kl_bwd(philox)", (M, NL), name, 10 * M * NL * sz, m, mn)
                # apply_weight: x (M, N=3, C) -> (M, C)
                N, C = 3, NL // 3
                w = torch.rand(M, N, device=dev, dtype=dt)
                out = torch.empty(M, C, device=dev, dtype=dt)
                gx, gw = torch.empty_like(mu), torch.empty_like(w)
                go = gz[:, :C].contiguous()

                def afwd():
                    L.check(lib.mlvae_apply_weight_fwd(L.ptr(mu), L.ptr(w), M, N, C, code, L.ptr(out), st))

                def abwd():
                    L.check(lib.mlvae_apply_weight_bwd(L.ptr(mu), L.ptr(w), L.ptr(go), M, N, C, code, L.ptr(gx), L.ptr(gw), st))
                m, mn = time_ms(afwd, None if big else flush)
                rec("apply_weight_fwd", (M, N, C), name, (M * N * C + M * N + M * C) * sz, m, mn)
                m, mn = time_ms(abwd, None if big else flush)
                rec("apply_weight_bwd", (M, N, C), name, (2 * M * N * C + 2 * M * N + M * C) * sz, m, mn)
                del ts, mu, lv, pmu, plv, gz, gk, z, kl, outs, w, out, gx, gw, go
                torch.cuda.empty_cache()

    if not args.only or "dense" in args.only:
        from ml_vae_b200 import dense
        for M, N in [(32000, 128), (1 << 20, 64), (1 << 20, 128)]:
            dy = torch.randn(M, N, device=dev).bfloat16()
            y = torch.randn(M, N, device=dev).bfloat16()
            m, mn = time_ms(lambda: dense._bwd_prep(dy, y), flush)
            rec("dense_bwd_prep(leaky+db)", (M, N), "bf16", 3 * M * N * 2, m, mn)
            m, mn = time_ms(lambda: dense._bwd_prep(dy, None), flush)
            rec("dense_bwd_prep(db)", (M, N), "bf16", M * N * 2, m, mn)

    if not args.only or "pcm" in args.only:
        for B, n in [(64, 80000), (512, 80000)]:
            blob = torch.randint(-32768, 32767, (B * n,), device=dev, dtype=torch.int16)
            offs = torch.arange(B, device=dev, dtype=torch.int64) * n
            lens = torch.full((B,), n, device=dev, dtype=torch.int32)
            out = torch.empty(B, n, device=dev)
            f = lambda: L.check(L.lib().mlvae_pcm_unpack(L.ptr(blob), 0, L.ptr(offs), L.ptr(lens), B, n, 1.0 / 32768.0, L.ptr(out),
                                                         L.stream_ptr()))
            m, mn = time_ms(f, flush)
            rec("pcm_unpack(int16->f32)", (B, n), "i16", B * n * 6, m, mn)

    if args.out:
        os.makedirs(os.path.dirname(os.path.join(ROOT, args.out)), exist_ok=True)
        json.dump({"peak_gbs": peak, "rows": rows}, open(os.path.join(ROOT, args.out), "w"), indent=1)


if __name__ == "__main__":
    main()
