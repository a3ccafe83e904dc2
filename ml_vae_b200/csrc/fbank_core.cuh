// Register-level pieces of the 400-point real FFT used by the fused fbank kernel.
// __host__ __device__ so tests/host/fbank_core_test.cu can run the very same code on
// the CPU against a float64 DFT (no GPU needed for the algebra).
//
// 400-point real FFT = 200-point complex FFT of z[m] = x[2m] + i x[2m+1] + split step.
// 200 = 25 x 8 (Cooley-Tukey, m = 8*n1 + n2, k = k1 + 25*k2):
//   pass A (per n2):  Y[n2][k1] = W200^(n2*k1) * sum_n1 z[8 n1 + n2] W25^(n1 k1)     25-point DFT = 5 x 5
//   pass B (per k1):  Z[k1 + 25 k2] = sum_n2 Y[n2][k1] W8^(n2 k2)                    radix-8
//   split:            X[k], X[200-k] from Z[k], Z[200-k]  ->  power = |X|^2
#pragma once
#include <cuda_runtime.h>

#ifndef MLVAE_HD
#define MLVAE_HD __host__ __device__ __forceinline__
#endif

namespace mlvae {

struct cpx {
    float re, im;
};
MLVAE_HD cpx cmul(cpx a, cpx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
MLVAE_HD cpx cadd(cpx a, cpx b) { return {a.re + b.re, a.im + b.im}; }
MLVAE_HD cpx csub(cpx a, cpx b) { return {a.re - b.re, a.im - b.im}; }

// In-place forward 5-point DFT (W = exp(-2 pi i / 5)) on x[0], x[s], ..., x[4s].
template <int S>
MLVAE_HD void dft5(cpx *x) {
    constexpr float C1 = 0.30901699437494742f;    // cos(2pi/5)
    constexpr float C2 = -0.80901699437494742f;   // cos(4pi/5)
    constexpr float S1 = 0.95105651629515357f;    // sin(2pi/5)
    constexpr float S2 = 0.58778525229247313f;    // sin(4pi/5)
    const cpx x0 = x[0], x1 = x[S], x2 = x[2 * S], x3 = x[3 * S], x4 = x[4 * S];
    const cpx t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    x[0] = {x0.re + t1.re + t2.re, x0.im + t1.im + t2.im};
    const cpx m1 = {x0.re + C1 * t1.re + C2 * t2.re, x0.im + C1 * t1.im + C2 * t2.im};
    const cpx m2 = {x0.re + C2 * t1.re + C1 * t2.re, x0.im + C2 * t1.im + C1 * t2.im};
    const cpx s1 = {S1 * t3.re + S2 * t4.re, S1 * t3.im + S2 * t4.im};
    const cpx s2 = {S2 * t3.re - S1 * t4.re, S2 * t3.im - S1 * t4.im};
    // y1 = m1 - i s1, y4 = m1 + i s1, y2 = m2 - i s2, y3 = m2 + i s2
    x[S] = {m1.re + s1.im, m1.im - s1.re};
    x[4 * S] = {m1.re - s1.im, m1.im + s1.re};
    x[2 * S] = {m2.re + s2.im, m2.im - s2.re};
    x[3 * S] = {m2.re - s2.im, m2.im + s2.re};
}

// 25-point forward DFT in registers.  in: a[n1], n1 = 5p + q.  out: a[k1] natural order.
// tw25[q*5 + r] = W25^(q r).
MLVAE_HD void dft25(cpx *a, const cpx *tw25) {
    // (a) for each q: 5-point DFT over p (stride 5) -> a[5r + q] = A_q[r]
#pragma unroll
    for (int q = 0; q < 5; ++q) dft5<5>(a + q);
    // (b) twiddle A_q[r] *= W25^(q r)
#pragma unroll
    for (int q = 1; q < 5; ++q)
#pragma unroll
        for (int r = 1; r < 5; ++r) a[5 * r + q] = cmul(a[5 * r + q], tw25[q * 5 + r]);
    // (c) for each r: 5-point DFT over q (stride 1) -> a[5r + s] = Y[r + 5 s]
#pragma unroll
    for (int r = 0; r < 5; ++r) dft5<1>(a + 5 * r);
    // (d) reorder to natural k1 = r + 5 s  (register renaming only, fully unrolled)
    cpx t[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) t[i] = a[i];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int s = 0; s < 5; ++s) a[r + 5 * s] = t[5 * r + s];
}

// In-place forward 8-point DFT, natural order in and out.
MLVAE_HD void dft8(cpx *x) {
    constexpr float H = 0.70710678118654752f;
    // stage 1: radix-2 over (j, j+4)
    cpx a0 = cadd(x[0], x[4]), a4 = csub(x[0], x[4]);
    cpx a1 = cadd(x[1], x[5]), a5 = csub(x[1], x[5]);
    cpx a2 = cadd(x[2], x[6]), a6 = csub(x[2], x[6]);
    cpx a3 = cadd(x[3], x[7]), a7 = csub(x[3], x[7]);
    // twiddles W8^j on the odd half: W8 = (1 - i)/sqrt2, W8^2 = -i, W8^3 = (-1 - i)/sqrt2
    a5 = {H * (a5.re + a5.im), H * (a5.im - a5.re)};
    a6 = {a6.im, -a6.re};
    a7 = {H * (a7.im - a7.re), -H * (a7.re + a7.im)};
    // even outputs: 4-point DFT of a0..a3 ; odd outputs: 4-point DFT of a4..a7
    cpx b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = csub(a1, a3);
    b3 = {b3.im, -b3.re};   // * -i
    x[0] = cadd(b0, b1); x[4] = csub(b0, b1); x[2] = cadd(b2, b3); x[6] = csub(b2, b3);
    cpx c0 = cadd(a4, a6), c2 = csub(a4, a6), c1 = cadd(a5, a7), c3 = csub(a5, a7);
    c3 = {c3.im, -c3.re};
    x[1] = cadd(c0, c1); x[5] = csub(c0, c1); x[3] = cadd(c2, c3); x[7] = csub(c2, c3);
}

// Split step of the real FFT for the pair (k, 200-k): Zk = Z[k], Zm = Z[200-k] (Z[200] := Z[0]),
// w = exp(-2 pi i k / 400).  Returns |X[k]|^2 and |X[200-k]|^2.
MLVAE_HD void split_power(cpx Zk, cpx Zm, cpx w, float &pk, float &pm) {
    const cpx E = {0.5f * (Zk.re + Zm.re), 0.5f * (Zk.im - Zm.im)};
    const cpx O = {0.5f * (Zk.im + Zm.im), -0.5f * (Zk.re - Zm.re)};   // -(i/2)(Zk - conj Zm)
    const cpx wo = cmul(w, O);
    const cpx xk = cadd(E, wo), xm = csub(E, wo);
    pk = xk.re * xk.re + xk.im * xk.im;
    pm = xm.re * xm.re + xm.im * xm.im;
}

}  // namespace mlvae
