"""Generate tests/golden/md_decode_cases.npz by running the REFERENCE's own decoder,
/root/reference/src/utils/decode_utils.py::decode_plvl_md_lbl_seqs_full (imported unmodified), on seeded inputs, and check
that oracle/decode_ref.py reproduces its integer outputs bit for bit.  Build container only (needs /root/reference):

    python oracle/gen_golden_decode.py

Stored per case: the model outputs the reference was given (logits, boundary probabilities, pi logits, prior, canonical
sequences, relative lengths, weight), the log-probability arrays its own pre-computation produced from them (same torch calls,
same `log()` helper: decode_utils.py:8-14, 421-438), and its three outputs padded with -1.  numpy here is >= 2 (NEP 50 promotion),
which is recorded in the file because two float32 spots of the reference depend on it (see oracle/decode_ref.py).

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = "/root/reference/src"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import decode_ref  # noqa: E402


def make_case(seed, B, T, N, Lmax, weight, extreme=False, tight=False):
    g = torch.Generator().manual_seed(seed)
    logits = 3.0 * torch.randn(B, T, N, generator=g)
    boundary_v = torch.rand(B, T, generator=g)
    pi_logits = 2.0 * torch.randn(B, T, 2, generator=g)
    prior = 0.05 + 0.9 * torch.rand(N, generator=g)
    if extreme:                                   # saturated probabilities: exercises the clamp of log() on both sides
        logits[:, ::3] *= 40.0
        boundary_v[:, ::4] = 0.0
        boundary_v[:, 1::4] = 1.0
        pi_logits[:, ::5] *= 60.0
    y = torch.randint(0, N, (B, Lmax), generator=g)
    t_abs = torch.randint(max(Lmax, T // 2), T + 1, (B,), generator=g)
    t_abs[0] = T
    l_abs = torch.randint(1, Lmax + 1, (B,), generator=g)
    l_abs[0] = Lmax
    if B > 1:
        l_abs[1] = 1                              # a single phoneme: only the l == 0 branch
    if tight and B > 2:
        t_abs[2] = int(l_abs[2])                  # T_i == L_i: one frame per phoneme, the only feasible path
    return dict(logits=logits, boundary_v=boundary_v, pi_logits=pi_logits, prior=prior, y=y,
                feat_lens=t_abs.float() / T, seq_lens=l_abs.float() / Lmax, weight=weight)


def main():
    if not os.path.isdir(REF_SRC):
        raise SystemExit("gen_golden_decode.py needs /root/reference (build container only)")
    sys.path.insert(0, REF_SRC)
    from utils import decode_utils as ref                      # the reference, unmodified

    cases = [make_case(1, 5, 40, 9, 6, 1.0), make_case(2, 4, 64, 12, 11, 0.7, extreme=True),
             make_case(3, 6, 33, 7, 8, 1.0, extreme=True, tight=True), make_case(4, 3, 120, 20, 17, 2.5)]
    out = {"numpy_major": np.array(int(np.__version__.split(".")[0])), "n_cases": np.array(len(cases))}
    for k, c in enumerate(cases):
        B, T, N = c["logits"].shape
        Lmax = c["y"].shape[1]
        preds = {"phn_recog_out": c["logits"], "boundary_v": c["boundary_v"], "pi_logits": c["pi_logits"]}
        bnd, fl, pl = ref.decode_plvl_md_lbl_seqs_full(preds, [f"utt{i}" for i in range(B)], c["feat_lens"], c["y"], c["seq_lens"],
                                                      c["prior"], weight=c["weight"])
        # the log-probabilities exactly as the reference's pre-computation builds them (decode_utils.py:421-438)
        po = torch.sigmoid(c["logits"])
        log_p_yx = ref.log(torch.stack([po, 1 - po], dim=3))
        log_p_y = ref.log(torch.stack([c["prior"], 1 - c["prior"]], dim=1))
        log_p_b = ref.log(torch.stack([c["boundary_v"], 1 - c["boundary_v"]], dim=2))
        log_p_pi = ref.log(torch.softmax(c["pi_logits"], dim=-1))
        t_abs = torch.round(c["feat_lens"] * T).int().numpy()
        l_abs = torch.round(c["seq_lens"] * Lmax).int().numpy()
        boundary = -np.ones((B, T), dtype=np.int32)
        frames = -np.ones((B, T), dtype=np.int32)
        phones = -np.ones((B, Lmax), dtype=np.int32)
        for i in range(B):
            boundary[i, :t_abs[i]] = bnd[i]
            frames[i, :t_abs[i]] = fl[i]
            phones[i, :l_abs[i]] = pl[i]
        # pin the restatement
        ob, of, op = decode_ref.decode_batch(log_p_yx, log_p_b, log_p_pi, log_p_y, c["y"].numpy(), t_abs, l_abs, c["weight"],
                                             numpy2=int(np.__version__.split(".")[0]) >= 2)
        for i in range(B):
            assert np.array_equal(ob[i], bnd[i]) and list(of[i]) == list(fl[i]) and list(op[i]) == list(pl[i]), (k, i)
        for name, v in dict(logits=c["logits"], boundary_v=c["boundary_v"], pi_logits=c["pi_logits"], prior=c["prior"], y=c["y"],
                            feat_lens=c["feat_lens"], seq_lens=c["seq_lens"]).items():
            out[f"c{k}.{name}"] = v.numpy()
        out[f"c{k}.weight"] = np.array(c["weight"])
        out[f"c{k}.log_p_yx"], out[f"c{k}.log_p_y"], out[f"c{k}.log_p_b"], out[f"c{k}.log_p_pi"] = log_p_yx, log_p_y, log_p_b, log_p_pi
        out[f"c{k}.t_abs"], out[f"c{k}.l_abs"] = t_abs, l_abs
        out[f"c{k}.boundary"], out[f"c{k}.frames"], out[f"c{k}.phones"] = boundary, frames, phones
        print(f"case {k}: B={B} T={T} N={N} Lmax={Lmax} weight={c['weight']}: reference == oracle "
              f"({int((frames == 1).sum())} mispronounced frames of {int((frames >= 0).sum())})")
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "md_decode_cases.npz"), **out)
    print("wrote", os.path.join(GOLD, "md_decode_cases.npz"))


if __name__ == "__main__":
    main()
