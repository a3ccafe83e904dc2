import numpy as np
import torch


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  (norm-wise relative error; b is the reference)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, rtol, what="", atol_scale=None):
    """elementwise |a-b| <= rtol * (|b| + scale), scale = atol_scale or max|b|: the stated
    relative tolerance applied with a floor so that values crossing zero are testable."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = float(b.abs().max()) if atol_scale is None else atol_scale
    bad = (a - b).abs() > rtol * (b.abs() + scale)
    nan_mismatch = torch.isnan(a) != torch.isnan(b)
    bad = (bad & ~torch.isnan(b)) | nan_mismatch
    if bad.any():
        i = int(torch.nonzero(bad.reshape(-1))[0])
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements outside rtol={rtol}; first at {i}: "
                             f"got {a.reshape(-1)[i].item()!r} want {b.reshape(-1)[i].item()!r}; "
                             f"max rel-to-max err {rel_err(a, b):.3e}")


FP32_RTOL = 1e-5      # BASELINE.json north_star: fp32 1e-5 rel
BF16_RTOL = 1e-2      # BASELINE.json north_star: bf16 1e-2 rel
