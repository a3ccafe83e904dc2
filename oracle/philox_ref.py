"""CPU oracle: Philox4x32-10 counter RNG and the eps ~ N(0,1) stream the CUDA
reparameterisation kernel draws from.  TEST INFRASTRUCTURE ONLY.

The reference draws eps with torch.randn_like (vanilla_vae.py:39) from the
global generator, which no other implementation can reproduce; the B200 path
replaces it by a counter-based stream keyed by (seed, offset, element index)
so that forward and backward regenerate the same eps without storing it.
This file is the host restatement of that stream (numpy, vectorised):

  Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy
  as 1, 2, 3" (SC'11); constants M0=0xD2511F53, M1=0xCD9E8D57,
  W0=0x9E3779B9, W1=0xBB67AE85.  Known-answer vectors from the Random123
  distribution (kat_vectors) are checked in tests/test_philox.py.

  Element i of the stream (flat, row-major index into the (M, L) latent):
    block   q = i // 4, lane r = i % 4
    counter = (q & 0xffffffff, q >> 32, offset & 0xffffffff, offset >> 32)
    key     = (seed & 0xffffffff, seed >> 32)
    u32[4]  = philox4x32_10(counter, key)
    Box-Muller on pairs: (u32[0], u32[1]) -> normals 0,1 ; (u32[2], u32[3]) -> 2,3
      u1    = fl32(fl32(a) + 1) * 2^-32               in (0, 1]     (float32 steps, like the kernel)
      theta = fl32(int32(b)) * fl32(pi * 2^-31)        in [-pi, pi)
      rad = sqrt(-2 ln u1);  n_even = rad * cos(theta);  n_odd = rad * sin(theta)
    The kernel evaluates ln/sqrt/sin/cos with the GPU's fast approximations (abs error
    ~1e-6); this file evaluates them in float64 from the same float32 u1/theta, so the two
    agree to ~2e-6 absolute -- parity tests that need identical eps materialise it on the
    device (mlvae_philox_normal) and feed that tensor to the oracle.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr (..., 4) uint32, key (..., 2) uint32 -> (..., 4) uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def philox_u32(seed: int, offset: int, n: int) -> np.ndarray:
    """First n uint32 of the element stream described in the module docstring."""
    nblk = (n + 3) // 4
    q = np.arange(nblk, dtype=np.uint64)
    ctr = np.stack([q & MASK, q >> np.uint64(32),
                    np.full(nblk, offset & 0xFFFFFFFF, np.uint64),
                    np.full(nblk, (offset >> 32) & 0xFFFFFFFF, np.uint64)], -1).astype(np.uint32)
    key = np.empty((nblk, 2), np.uint32)
    key[:, 0] = seed & 0xFFFFFFFF
    key[:, 1] = (seed >> 32) & 0xFFFFFFFF
    return philox4x32_10(ctr, key).reshape(-1)[:n]


def philox_normal(seed: int, offset: int, n: int, dtype=np.float32) -> np.ndarray:
    """eps stream, computed in float64 and rounded once to ``dtype``."""
    nblk = (n + 3) // 4
    u = philox_u32(seed, offset, nblk * 4).reshape(nblk, 2, 2)
    u1 = ((u[..., 0].astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -32)).astype(np.float64)
    theta = (u[..., 1].view(np.int32).astype(np.float32) * np.float32(np.pi * 2.0 ** -31)).astype(np.float64)
    rad = np.sqrt(-2.0 * np.log(u1))
    out = np.stack([rad * np.cos(theta), rad * np.sin(theta)], -1)
    return out.reshape(-1)[:n].astype(dtype)


def philox_normal_v2(seed: int, offset: int, n: int, dtype=np.float32) -> np.ndarray:
    """eps stream of the bf16 kernels (csrc/philox.cuh "v2"): EIGHT normals per Philox block from 16-bit uniforms.
    Element i: block q = i // 8 (counter (q, offset), key seed), word j = (i % 8) // 2 of the block gives the pair (2j, 2j+1):
      u1 = (lo16 + 1) * 2^-16 in (0, 1];  theta = (hi16 - 32768) * fl32(pi / 32768);  r = sqrt(-2 ln u1)
      n_even = r cos(theta), n_odd = r sin(theta)
    evaluated in float64 from the same float32 u1 / theta as the kernel (which uses the GPU's fast approximations)."""
    nblk = (n + 7) // 8
    q = np.arange(nblk, dtype=np.uint64)
    ctr = np.stack([q & MASK, q >> np.uint64(32),
                    np.full(nblk, offset & 0xFFFFFFFF, np.uint64),
                    np.full(nblk, (offset >> 32) & 0xFFFFFFFF, np.uint64)], -1).astype(np.uint32)
    key = np.empty((nblk, 2), np.uint32)
    key[:, 0] = seed & 0xFFFFFFFF
    key[:, 1] = (seed >> 32) & 0xFFFFFFFF
    w = philox4x32_10(ctr, key)                                   # (nblk, 4)
    lo = (w & np.uint32(0xFFFF)).astype(np.float32)
    hi = (w >> np.uint32(16)).astype(np.int64)
    u1 = ((lo + np.float32(1.0)) * np.float32(2.0 ** -16)).astype(np.float64)
    theta = ((hi - 32768).astype(np.float32) * np.float32(np.pi / 32768.0)).astype(np.float64)
    rad = np.sqrt(-2.0 * np.log(u1))
    out = np.stack([rad * np.cos(theta), rad * np.sin(theta)], -1)   # (nblk, 4, 2)
    return out.reshape(-1)[:n].astype(dtype)


def dropout_keep_mask(seed: int, offset: int, n: int, p: float) -> np.ndarray:
    """Keep mask (bool, n) of the inter-layer LSTM dropout (decoder.py:14-15, dropout=rnn_dropout) as the CUDA kernel
    csrc/dropout.cu draws it: element i uses the 16-bit lane i % 8 (word (i % 8) // 2, low half first) of the Philox
    block with counter (i // 8, offset) and key seed; keep iff u16 >= round(p * 65536).  The kept values are scaled by
    float32(1 / (1 - p)), like torch.nn.functional.dropout."""
    nblk = (n + 7) // 8
    q = np.arange(nblk, dtype=np.uint64)
    ctr = np.stack([q & MASK, q >> np.uint64(32),
                    np.full(nblk, offset & 0xFFFFFFFF, np.uint64),
                    np.full(nblk, (offset >> 32) & 0xFFFFFFFF, np.uint64)], -1).astype(np.uint32)
    key = np.empty((nblk, 2), np.uint32)
    key[:, 0] = seed & 0xFFFFFFFF
    key[:, 1] = (seed >> 32) & 0xFFFFFFFF
    w = philox4x32_10(ctr, key)                                   # (nblk, 4)
    u16 = np.stack([w & np.uint32(0xFFFF), w >> np.uint32(16)], -1).reshape(nblk * 8)[:n]
    thresh = int(np.rint(np.float32(p) * np.float32(65536.0)))
    return u16 >= thresh
