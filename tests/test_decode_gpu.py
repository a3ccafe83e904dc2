"""GPU parity: the joint boundary / mispronunciation decoder (csrc/md_decode.cu through the C ABI) against
  * the golden vectors produced by the reference's own decode_plvl_md_lbl_seqs_full (utils/decode_utils.py:374-565;
    oracle/gen_golden_decode.py -> tests/golden/md_decode_cases.npz), and
  * the CPU restatement oracle/decode_ref.py on seeded inputs at BASELINE-sized shapes.
Integer outputs: bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import decode_ref

pytestmark = pytest.mark.gpu


def _golden():
    return np.load(os.path.join(GOLDEN, "md_decode_cases.npz"))


def _check_against_lists(boundary, frames, phones, status, ob, of, op, t_abs, l_abs):
    boundary, frames, phones, status = (x.cpu().numpy() for x in (boundary, frames, phones, status))
    assert (status == 0).all(), status
    for i in range(len(t_abs)):
        Ti, Li = int(t_abs[i]), int(l_abs[i])
        assert np.array_equal(boundary[i, :Ti], np.asarray(ob[i])), f"boundary of utterance {i}"
        assert np.array_equal(frames[i, :Ti], np.asarray(of[i])), f"frame labels of utterance {i}"
        assert np.array_equal(phones[i, :Li], np.asarray(op[i])), f"phoneme labels of utterance {i}"
        assert (boundary[i, Ti:] == -1).all() and (frames[i, Ti:] == -1).all() and (phones[i, Li:] == -1).all()


def test_kernel_matches_reference_golden_on_identical_logs(cuda):
    from ml_vae_b200.utils import decode_utils as du
    g = _golden()
    for k in range(int(g["n_cases"])):
        c = lambda n: g[f"c{k}.{n}"]
        out = du.decode_from_logs(c("log_p_yx"), c("log_p_b"), c("log_p_pi"), c("log_p_y"), c("y"), c("t_abs"), c("l_abs"),
                                  weight=float(c("weight")), numpy2=int(g["numpy_major"]) >= 2, device=cuda)
        B = c("y").shape[0]
        ob = [c("boundary")[i, :c("t_abs")[i]] for i in range(B)]
        of = [c("frames")[i, :c("t_abs")[i]] for i in range(B)]
        op = [c("phones")[i, :c("l_abs")[i]] for i in range(B)]
        _check_against_lists(*out, ob, of, op, c("t_abs"), c("l_abs"))


def test_drop_in_function_matches_reference_golden(cuda):
    """Same arguments as the reference function, model outputs on the GPU, pre-computation call for call (log on the CPU)."""
    from ml_vae_b200.utils.decode_utils import decode_plvl_md_lbl_seqs_full
    g = _golden()
    if int(g["numpy_major"]) != int(np.__version__.split(".")[0]):
        pytest.skip("golden vectors were generated under another numpy major version (promotion rules differ)")
    for k in range(int(g["n_cases"])):
        c = lambda n: g[f"c{k}.{n}"]
        preds = {"phn_recog_out": torch.from_numpy(c("logits")).to(cuda), "boundary_v": torch.from_numpy(c("boundary_v")).to(cuda),
                 "pi_logits": torch.from_numpy(c("pi_logits")).to(cuda)}
        B = c("y").shape[0]
        bnd, fl, pl = decode_plvl_md_lbl_seqs_full(preds, [f"u{i}" for i in range(B)], torch.from_numpy(c("feat_lens")).to(cuda),
                                                   torch.from_numpy(c("y")).to(cuda), torch.from_numpy(c("seq_lens")).to(cuda),
                                                   torch.from_numpy(c("prior")), weight=float(c("weight")))
        mism = 0
        for i in range(B):
            Ti, Li = int(c("t_abs")[i]), int(c("l_abs")[i])
            assert len(bnd[i]) == Ti and len(fl[i]) == Ti and len(pl[i]) == Li and int(bnd[i].sum()) == Li
            mism += int(not (np.array_equal(bnd[i], c("boundary")[i, :Ti]) and fl[i] == list(c("frames")[i, :Ti]) and pl[i] == list(c("phones")[i, :Li])))
        # sigmoid / softmax on the GPU may differ from the CPU's in the last ulp: identical labels are expected, a flip would
        # need an exact near-tie on the winning path
        assert mism == 0, f"case {k}: {mism} utterances differ from the reference"


def _random_logs(seed, B, T, N, Lmax, neg_inf=False):
    rng = np.random.default_rng(seed)
    p = 1.0 / (1.0 + np.exp(-3.0 * rng.standard_normal((B, T, N)))).astype(np.float32)
    log_p_yx = decode_ref.ref_log(np.stack([p, 1 - p], axis=3))
    bv = rng.random((B, T)).astype(np.float32)
    log_p_b = decode_ref.ref_log(np.stack([bv, 1 - bv], axis=2))
    z = (2.0 * rng.standard_normal((B, T, 2))).astype(np.float32)
    e = np.exp(z - z.max(-1, keepdims=True))
    log_p_pi = decode_ref.ref_log((e / e.sum(-1, keepdims=True)).astype(np.float32))
    prior = (0.05 + 0.9 * rng.random(N)).astype(np.float32)
    log_p_y = decode_ref.ref_log(np.stack([prior, 1 - prior], axis=1))
    if neg_inf:                                  # forbidden boundaries at some frames: -inf must propagate like numpy's
        log_p_b[:, 5::7, 1] = -np.inf
    y = rng.integers(0, N, (B, Lmax)).astype(np.int32)
    t_abs = rng.integers(max(Lmax, T // 2), T + 1, B).astype(np.int32)
    l_abs = rng.integers(1, Lmax + 1, B).astype(np.int32)
    t_abs[0], l_abs[0] = T, Lmax
    return log_p_yx, log_p_b, log_p_pi, log_p_y, y, t_abs, l_abs


@pytest.mark.parametrize("B,T,N,Lmax,weight,numpy2,neg_inf", [
    (16, 500, 42, 60, 1.0, True, False),          # BASELINE configs[1]-sized batch: 5 s utterances, ARPAbet-sized inventory
    (8, 500, 42, 33, 0.37, False, False),         # numpy 1.x promotion (float64 weight product / first cell)
    (8, 300, 40, 47, 2.0, True, True),            # -inf transitions
    (3, 2000, 42, 100, 1.0, True, False),         # configs[3]-sized: 20 s utterances, back pointers still on chip (203 KB)
    (2, 2000, 42, 128, 1.0, True, False),         # T x Lmax = 256 KB -> back pointers in the caller's global workspace
    (4, 64, 5, 64, 1.0, True, False),             # Lmax == T possible; two full warps of phonemes
    (2, 1, 3, 1, 1.0, True, False),               # a single frame
])
def test_kernel_matches_oracle(cuda, B, T, N, Lmax, weight, numpy2, neg_inf):
    from ml_vae_b200.utils import decode_utils as du
    args = _random_logs(B * 1000 + T + Lmax, B, T, N, Lmax, neg_inf)
    if neg_inf:
        # make sure every utterance still has a feasible path: at least L-1 frames with a finite boundary score remain
        assert all(int(np.isfinite(args[1][i, 1:args[5][i], 1]).sum()) >= int(args[6][i]) - 1 for i in range(B))
    ob, of, op = decode_ref.decode_batch(*args, weight=weight, numpy2=numpy2)
    out = du.decode_from_logs(*args, weight=weight, numpy2=numpy2, device=cuda)
    _check_against_lists(*out, ob, of, op, args[5], args[6])


def test_infeasible_lengths_raise_like_the_reference(cuda):
    """More phonemes than frames: the reference dies in `assert l == t == 0` (decode_utils.py:536); so does the drop-in."""
    from ml_vae_b200.utils import decode_utils as du
    from ml_vae_b200.utils.decode_utils import decode_plvl_md_lbl_seqs_full
    args = list(_random_logs(7, 3, 12, 6, 10))
    args[5][:] = [12, 4, 12]
    args[6][:] = [10, 9, 3]                       # utterance 1: 9 phonemes, 4 frames
    status = du.decode_from_logs(*args, device=cuda)[3].cpu().numpy()
    assert list(status) == [0, 1, 0]
    bad_y = args[4].copy(); bad_y[2, 1] = 6       # phoneme index outside the inventory
    assert int(du.decode_from_logs(args[0], args[1], args[2], args[3], bad_y, args[5], args[6], device=cuda)[3][2]) == 2
    preds = {"phn_recog_out": torch.randn(2, 6, 5, device=cuda), "boundary_v": torch.rand(2, 6, device=cuda), "pi_logits": torch.randn(2, 6, 2, device=cuda)}
    with pytest.raises(AssertionError):
        decode_plvl_md_lbl_seqs_full(preds, ["a", "b"], torch.tensor([1.0, 0.5]), torch.randint(0, 5, (2, 4)), torch.tensor([1.0, 1.0]), torch.rand(5))


def test_cpu_model_outputs_are_refused(lib_built):
    from ml_vae_b200._lib import MlvaeError
    from ml_vae_b200.utils.decode_utils import decode_plvl_md_lbl_seqs_full
    preds = {"phn_recog_out": torch.randn(1, 6, 5), "boundary_v": torch.rand(1, 6), "pi_logits": torch.randn(1, 6, 2)}
    with pytest.raises(MlvaeError):
        decode_plvl_md_lbl_seqs_full(preds, ["a"], torch.tensor([1.0]), torch.randint(0, 5, (1, 3)), torch.tensor([1.0]), torch.rand(5))
