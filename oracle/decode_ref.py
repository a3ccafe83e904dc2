"""CPU restatement of the reference's joint boundary / mispronunciation decoder (SURVEY.md section 8 f4 "later"):

    utils/decode_utils.py:374-565   decode_plvl_md_lbl_seqs_full   (identical DP: decode_plvl_md_lbl_seqs_full_non_par, :191-371)
    utils/decode_utils.py:8-14      log()  (clamp [0, 1e-5) to 1e-5, torch.log on the CPU, -> numpy float32)

TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this file; the product
(ml_vae_b200/) never does.  PINNED: oracle/gen_golden_decode.py runs the reference function itself (imported unmodified from
/root/reference/src/utils/decode_utils.py) on seeded inputs and stores inputs + outputs in tests/golden/md_decode_cases.npz;
tests/test_oracle_golden.py checks this restatement against those integer outputs bit for bit.

The algorithm (decode_utils.py:440-548).  Per utterance i with T_i frames and the canonical phoneme sequence y (L_i entries), a
Viterbi search over states (l, t, beta): phoneme index l, frame t, beta = 0 correct / 1 mispronounced.

    emit(t, l, s)   = log_p_yx[t, y_l, s] - log_p_y[y_l, s]
    dp[0, 0, s]     = weight * log_p_pi[0, s] + log_p_yx[0, y_0, s] - log_p_y[y_0, s]                        (:452-453)
    dp[0, t, s]     = dp[0, t-1, s] + log_p_b[t, 0] + emit                        path 0                       (:458-468)
    dp[l, t, s]     = max over  hold          dp[l,   t-1, s] + log_p_b[t, 0] + emit                            (:470-500)
                                from correct  dp[l-1, t-1, 0] + log_p_b[t, 1] + weight * log_p_pi[t, s] + emit
                                from incorr.  dp[l-1, t-1, 1] + log_p_b[t, 1] + weight * log_p_pi[t, s] + emit
                      path = np.argmax of that list (FIRST maximum wins)
    every sum is evaluated left to right exactly as written in the reference.
    backtracking (:503-536) from (L_i - 1, T_i - 1), beta = 0 iff dp[.., 0] > dp[.., 1] (strict).

Arithmetic types.  dp_value is a float64 numpy array and the log-probabilities are float32 numpy scalars, so every sum that starts
from a dp_value entry is a float64 sum of exactly-widened float32 terms under every numpy version.  Two spots depend on numpy's
promotion rules: `weight * log_p_pi[...]` (python float x np.float32) and the all-float32 initial row.  numpy >= 2 (NEP 50, what
this container runs and what the golden vectors were generated with) keeps both in float32; numpy 1.x (the reference's era,
value-based casting) evaluates them in float64.  ``numpy2`` selects which; with weight == 1.0 only the initial row differs.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-5


def ref_log(x: np.ndarray) -> np.ndarray:
    """decode_utils.py:8-14 in numpy float32 (np.log vs torch.log may differ in the last ulp: tests feed identical log inputs)."""
    r = np.array(x, dtype=np.float32, copy=True)
    r[(r >= 0) & (r < EPS)] = EPS
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.log(r)


def decode_one(log_p_yx, log_p_b, log_p_pi, log_p_y, y, weight=1.0, numpy2=True):
    """One utterance.  log_p_yx (T, N, 2), log_p_b (T, 2), log_p_pi (T, 2), log_p_y (N, 2) float32; y (L,) int.
    -> (boundary (T,) int, frame labels (T,) list, phoneme labels (L,) list)   decode_utils.py:440-548"""
    T, L = log_p_yx.shape[0], len(y)
    if L < 1 or T < 1 or L > T:
        raise AssertionError(f"l = {L}, t = {T}")          # the reference fails its final assert (or indexes out of range)
    f32, f64 = np.float32, np.float64
    lyx = log_p_yx[:, y, :].astype(f32)                      # (T, L, 2)
    ly = log_p_y[y, :].astype(f32)                           # (L, 2)
    lb, lpi = log_p_b.astype(f32), log_p_pi.astype(f32)
    if numpy2:
        wpi = (f32(weight) * lpi).astype(f64)                # float32 product, widened when added to a float64 sum
        init = ((f32(weight) * lpi[0] + lyx[0, 0]) - ly[0]).astype(f64)      # all-float32 chain (:452-453)
    else:
        wpi = f64(weight) * lpi.astype(f64)
        init = (wpi[0] + lyx[0, 0].astype(f64)) - ly[0].astype(f64)
    lyx64, ly64, lb64 = lyx.astype(f64), ly.astype(f64), lb.astype(f64)
    dp = np.full((L, 2), -np.inf, dtype=f64)
    dp[0] = init
    path = np.zeros((L, T, 2), dtype=np.int8)               # reference: -1 at t = 0 (never read)
    with np.errstate(invalid="ignore"):
        for t in range(1, T):
            prev = dp
            hold = ((prev + lb64[t, 0]) + lyx64[t]) - ly64                                     # (L, 2)
            cur = hold.copy()
            pth = np.zeros((L, 2), dtype=np.int8)
            if L > 1:
                fc = (((prev[:-1, 0:1] + lb64[t, 1]) + wpi[t][None, :]) + lyx64[t, 1:]) - ly64[1:]     # from dp[l-1, t-1, 0]
                fi = (((prev[:-1, 1:2] + lb64[t, 1]) + wpi[t][None, :]) + lyx64[t, 1:]) - ly64[1:]     # from dp[l-1, t-1, 1]
                h = hold[1:]
                best = h.copy()
                p = np.zeros_like(pth[1:])
                m = fc > best                                # np.argmax: first maximum wins -> strict comparisons in list order
                best = np.where(m, fc, best); p = np.where(m, 1, p)
                m = fi > best
                best = np.where(m, fi, best); p = np.where(m, 2, p)
                cur[1:] = best
                pth[1:] = p
            dp = cur
            path[:, t, :] = pth
    l, t = L - 1, T - 1
    bidx, fl, pl = [], [], []
    beta = 0 if dp[l, 0] > dp[l, 1] else 1
    fl.append(beta); pl.append(beta)
    while t > 0:
        pz = path[l, t, beta]
        if pz == 1:
            l -= 1; bidx.append(t); fl.append(0); pl.append(0); beta = 0
        elif pz == 2:
            l -= 1; bidx.append(t); fl.append(1); pl.append(1); beta = 1
        else:
            fl.append(fl[-1])
        t -= 1
        if l < 0:
            raise AssertionError("l = -1")
    bidx.append(t)
    if not (l == 0 and t == 0):
        raise AssertionError(f"l = {l}, t = {t}")
    fl.reverse(); pl.reverse()
    boundary = np.zeros(T, dtype=np.int64)
    boundary[bidx] = 1
    assert boundary.sum() == L
    return boundary, fl, pl


def decode_batch(log_p_yx, log_p_b, log_p_pi, log_p_y, y, feat_lens, seq_lens, weight=1.0, numpy2=True):
    """Absolute lengths in, lists out (decode_utils.py:550-555)."""
    bs, fls, pls = [], [], []
    for i in range(len(feat_lens)):
        Ti, Li = int(feat_lens[i]), int(seq_lens[i])
        b, f, p = decode_one(log_p_yx[i, :Ti], log_p_b[i, :Ti], log_p_pi[i, :Ti], log_p_y, np.asarray(y[i, :Li]), weight, numpy2)
        bs.append(b); fls.append(f); pls.append(p)
    return bs, fls, pls
