// Fused acoustic front-end for sm_100a: framing + Hamming window + 400-point real FFT +
// power + triangular mel bank + 10*log10 + per-utterance top_db floor + delta/delta-delta.
// Replaces speechbrain.lobes.features.Fbank as declared at config/run.yaml:39-44 and
// called at utils/data_io.py:197-201 (the arithmetic itself is SpeechBrain's; restated in
// oracle/fbank_ref.py).
//
// Two launches:
//   logmel_kernel  one CTA = 32 consecutive frames of one utterance (lane == frame, so all
//                  shared-memory traffic is conflict-free); audio span staged once in shared
//                  memory with coalesced loads; FFT passes held in registers (25-point DFT per
//                  thread, then radix-8 + split per thread); power -> sparse mel -> dB; tile
//                  written coalesced to a float32 scratch; CTA max -> one atomicMax/utterance.
//   finish_kernel  floor at (utterance max - 80 dB), regression deltas with replicate padding,
//                  Kaldi-length truncation, zero fill, dtype conversion, coalesced stores.
// Roofline: HBM for bytes (4*hop + D*s per frame) but the FFT makes the first kernel
// FP32/shared-memory bound; DESIGN.md states the measured split.
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "fbank_core.cuh"

namespace mlvae {
namespace {

constexpr int kNfft = 400;
constexpr int kBins = 201;
constexpr int kTile = 32;            // frames per CTA (lane == frame)
constexpr int kWarps = 8;            // == radix-8 factor: warp w owns residue n2 = w in pass A
constexpr int kThreadsFb = kWarps * 32;
constexpr float kAmin = 1e-10f;
constexpr float kTopDb = 80.f;
constexpr float kTenLog10Of2 = 3.0102999566398120f;   // 10*log10(x) = this * log2(x)

__constant__ float c_win[kNfft];
__constant__ float2 c_win2[kNfft / 2];   // same window as (w[2m], w[2m+1]) pairs for the fast kernel
__constant__ cpx c_tw25[25];         // W25^(q r)
__constant__ cpx c_tw200[8 * 25];    // W200^(n2 k1)
__constant__ cpx c_tw400[kBins];     // exp(-2 pi i k / 400)

__device__ __forceinline__ unsigned int f2ord(float f) {   // order-preserving float -> uint
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct MelTables {          // device pointers
    const int *lo;          // first bin of filter m
    const int *cnt;         // number of consecutive bins
    const int *woff;        // offset of its weights in w
    const float *w;
    int nnz;
};

// max that propagates NaN like torch.max / torch.clamp / amax do (fmaxf would drop it)
__device__ __forceinline__ float pmax(float a, float b) {     // branch free: two selects around fmaxf
    const float m = fmaxf(a, b);
    const float n = (b != b) ? b : m;
    return (a != a) ? a : n;
}

__device__ __forceinline__ int skew(int p, int s) { return s > 0 ? p + (p >> s) : p; }

// ------------------------------------------------------------------ pass 1 --
__global__ void __launch_bounds__(kThreadsFb, 2)
logmel_kernel(const float *__restrict__ wav, const int32_t *__restrict__ wav_len, int64_t n_max, int64_t n_stride,
              int hop, int skew_shift, int n_mels, MelTables mel, float *__restrict__ logmel, int t_full_max,
              unsigned int *__restrict__ max_buf, int vec4_ok) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kTile;
    const int64_t len_b = wav_len ? (int64_t)wav_len[b] : n_max;
    const int t_full = 1 + (int)(len_b / hop);
    if (t0 >= t_full) return;

    const int span = (kTile - 1) * hop + kNfft;
    const int span_sk = skew(span, skew_shift) + 1;
    float2 *s_Y = reinterpret_cast<float2 *>(smem_raw);                       // [200][32]
    float *s_P = reinterpret_cast<float *>(smem_raw + 200 * 32 * 8);          // [201][32]
    float *s_audio = s_P + kBins * 32;                                        // [span_sk]
    float *s_mw = s_audio + ((span_sk + 3) & ~3);                             // [nnz]
    int *s_lo = reinterpret_cast<int *>(s_mw + mel.nnz);
    int *s_cnt = s_lo + n_mels;
    int *s_woff = s_cnt + n_mels;
    float *s_O = reinterpret_cast<float *>(smem_raw);                         // aliases s_Y after pass B

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- stage the audio span [s0, s0 + span) with zero fill outside [0, len_b) ----
    const float *row = wav + (int64_t)b * n_stride;
    const int64_t s0 = (int64_t)t0 * hop - kNfft / 2;
    if (vec4_ok) {
        for (int p = tid * 4; p < span; p += kThreadsFb * 4) {
            const int64_t g = s0 + p;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g >= 0 && g + 3 < len_b) v = __ldg(reinterpret_cast<const float4 *>(row + g));
            else {
                if (g >= 0 && g < len_b) v.x = __ldg(row + g);
                if (g + 1 >= 0 && g + 1 < len_b) v.y = __ldg(row + g + 1);
                if (g + 2 >= 0 && g + 2 < len_b) v.z = __ldg(row + g + 2);
                if (g + 3 >= 0 && g + 3 < len_b) v.w = __ldg(row + g + 3);
            }
            s_audio[skew(p, skew_shift)] = v.x;
            if (p + 1 < span) s_audio[skew(p + 1, skew_shift)] = v.y;
            if (p + 2 < span) s_audio[skew(p + 2, skew_shift)] = v.z;
            if (p + 3 < span) s_audio[skew(p + 3, skew_shift)] = v.w;
        }
    } else {
        for (int p = tid; p < span; p += kThreadsFb) {
            const int64_t g = s0 + p;
            s_audio[skew(p, skew_shift)] = (g >= 0 && g < len_b) ? __ldg(row + g) : 0.f;
        }
    }
    for (int i = tid; i < mel.nnz; i += kThreadsFb) s_mw[i] = mel.w[i];
    for (int i = tid; i < n_mels; i += kThreadsFb) {
        s_lo[i] = mel.lo[i]; s_cnt[i] = mel.cnt[i]; s_woff[i] = mel.woff[i];
    }
    __syncthreads();

    // ---- pass A: warp = residue n2, lane = frame: windowed 25-point DFT + W200 twiddle ----
    {
        const int n2 = warp;
        cpx a[25];
        const int base = lane * hop + 2 * n2;
#pragma unroll
        for (int n1 = 0; n1 < 25; ++n1) {
            const int p = base + 16 * n1;
            a[n1].re = s_audio[skew(p, skew_shift)] * c_win[16 * n1 + 2 * n2];
            a[n1].im = s_audio[skew(p + 1, skew_shift)] * c_win[16 * n1 + 2 * n2 + 1];
        }
        dft25(a, c_tw25);
        if (n2 != 0) {
#pragma unroll
            for (int k1 = 1; k1 < 25; ++k1) a[k1] = cmul(a[k1], c_tw200[n2 * 25 + k1]);
        }
#pragma unroll
        for (int k1 = 0; k1 < 25; ++k1) s_Y[(n2 * 25 + k1) * 32 + lane] = make_float2(a[k1].re, a[k1].im);
    }
    __syncthreads();

    // ---- pass B: 13 items (k1 = j and 25 - j) per frame: radix-8, split, power ----
    for (int j = warp; j < 13; j += kWarps) {
        cpx y[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
            const float2 v = s_Y[(n2 * 25 + j) * 32 + lane];
            y[n2] = {v.x, v.y};
        }
        dft8(y);
        if (j == 0) {
            float pk, pm;
            split_power(y[0], y[0], c_tw400[0], pk, pm);
            s_P[0 * 32 + lane] = pk; s_P[200 * 32 + lane] = pm;
#pragma unroll
            for (int k2 = 1; k2 < 4; ++k2) {
                split_power(y[k2], y[8 - k2], c_tw400[25 * k2], pk, pm);
                s_P[(25 * k2) * 32 + lane] = pk; s_P[(200 - 25 * k2) * 32 + lane] = pm;
            }
            split_power(y[4], y[4], c_tw400[100], pk, pm);
            s_P[100 * 32 + lane] = pk;
        } else {
            cpx y2[8];
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) {
                const float2 v = s_Y[(n2 * 25 + (25 - j)) * 32 + lane];
                y2[n2] = {v.x, v.y};
            }
            dft8(y2);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const int k = j + 25 * k2;
                float pk, pm;
                split_power(y[k2], y2[7 - k2], c_tw400[k], pk, pm);
                s_P[k * 32 + lane] = pk; s_P[(200 - k) * 32 + lane] = pm;
            }
        }
    }
    __syncthreads();

    // ---- mel bank + dB: warp = filter (round robin), lane = frame ----
    const bool frame_ok = (t0 + lane) < t_full;
    float vmax = -INFINITY;
    const int ostride = n_mels + 1;
    for (int m = warp; m < n_mels; m += kWarps) {
        const int lo = s_lo[m], cnt = s_cnt[m];
        const float *wp = s_mw + s_woff[m];
        float acc = 0.f;
        for (int i = 0; i < cnt; ++i) acc = fmaf(s_P[(lo + i) * 32 + lane], wp[i], acc);
        const float db = kTenLog10Of2 * __log2f(pmax(acc, kAmin));
        s_O[lane * ostride + m] = db;
        if (frame_ok) vmax = pmax(vmax, db);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = pmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    __shared__ float s_wmax[kWarps];
    if (lane == 0) s_wmax[warp] = vmax;
    __syncthreads();
    if (tid == 0) {
        float m = s_wmax[0];
#pragma unroll
        for (int i = 1; i < kWarps; ++i) m = pmax(m, s_wmax[i]);
        atomicMax(max_buf + b, f2ord(m));
    }
    // ---- coalesced tile store: frames [t0, t0 + nf) x n_mels are contiguous in the scratch ----
    const int nf = min(kTile, t_full - t0);
    float *dst = logmel + ((int64_t)b * t_full_max + t0) * n_mels;
    for (int f = warp; f < nf; f += kWarps)
        for (int m = lane; m < n_mels; m += 32) dst[f * n_mels + m] = s_O[f * ostride + m];
}

// ------------------------------------------------------------- pass 1, fast --
// Specialisation for hop % 16 == 0 (160 = 10 ms, 320 = 20 ms at 16 kHz), same algorithm, ~2x fewer instructions:
//   * staging with 8-byte cp.async (zero fill at the utterance edges) into a span skewed by 2 floats per hop, so
//     that lane == frame reads are conflict-free 64-bit LDS with COMPILE-TIME offsets (no address arithmetic);
//   * the 1/2 factors of the real-FFT split are dropped (power comes out x4, undone by one multiply per mel bin);
//   * mel bank unrolled by 4 with float4 weights (zero padded), 3.25 instead of 15 instructions per term.
template <int HOP>
struct FastCfg {
    static constexpr int kSpan = (kTile - 1) * HOP + kNfft;
    static constexpr int kAudioFloats = ((kSpan + 2 * (kSpan / HOP + 1) + 3) & ~3);
};

MLVAE_HD void split_power_x4(cpx Zk, cpx Zm, cpx w, float &pk, float &pm) {
    const cpx E = {Zk.re + Zm.re, Zk.im - Zm.im};
    const cpx O = {Zk.im + Zm.im, Zm.re - Zk.re};
    const cpx wo = cmul(w, O);
    const cpx xk = cadd(E, wo), xm = csub(E, wo);
    pk = xk.re * xk.re + xk.im * xk.im;
    pm = xm.re * xm.re + xm.im * xm.im;
}

struct MelTables4 {         // device pointers; weights zero padded to groups of 4 per filter
    const int *lo;          // first bin of filter m
    const int *cnt4;        // number of 4-bin groups
    const int *woff4;       // offset (in float4) of its weights
    const float4 *w4;
    int n4;                 // total float4
};

template <int HOP>
__global__ void __launch_bounds__(kThreadsFb, 2)
logmel_fast_kernel(const float *__restrict__ wav, const int32_t *__restrict__ wav_len, int64_t n_max, int64_t n_stride,
                   int n_mels, MelTables4 mel, float *__restrict__ logmel, int t_full_max, unsigned int *__restrict__ max_buf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Cfg = FastCfg<HOP>;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kTile;
    const int64_t len_b = wav_len ? (int64_t)wav_len[b] : n_max;
    const int t_full = 1 + (int)(len_b / HOP);
    if (t0 >= t_full) return;

    float2 *s_Y = reinterpret_cast<float2 *>(smem_raw);                       // [200][32]
    float *s_P = reinterpret_cast<float *>(smem_raw + 200 * 32 * 8);          // [204][32] (3 zero rows of slack)
    float *s_audio = s_P + (kBins + 3) * 32;                                  // skewed span
    float4 *s_w4 = reinterpret_cast<float4 *>(s_audio + Cfg::kAudioFloats);   // [n4]
    int *s_lo = reinterpret_cast<int *>(s_w4 + mel.n4);
    int *s_cnt4 = s_lo + n_mels;
    int *s_woff4 = s_cnt4 + n_mels;
    float *s_O = reinterpret_cast<float *>(smem_raw);                         // aliases s_Y after pass B
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- stage the audio span with 8-byte cp.async; pair (g, g+1): s0 is even, so a pair never straddles sample 0 ----
    {
        const float *row = wav + (int64_t)b * n_stride;
        const int64_t s0 = (int64_t)t0 * HOP - kNfft / 2;
        for (int q = tid; q < Cfg::kSpan / 2; q += kThreadsFb) {
            const int p = 2 * q;
            const int64_t g = s0 + p;
            uint32_t bytes = 0;
            if (g >= 0 && g < len_b) bytes = (g + 1 < len_b) ? 8u : 4u;
            const float *src = row + (bytes ? g : 0);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_audio + p + 2 * (p / HOP));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int i = tid; i < mel.n4; i += kThreadsFb) s_w4[i] = mel.w4[i];
    for (int i = tid; i < n_mels; i += kThreadsFb) {
        s_lo[i] = mel.lo[i]; s_cnt4[i] = mel.cnt4[i]; s_woff4[i] = mel.woff4[i];
    }
    if (tid < 96) s_P[kBins * 32 + tid] = 0.f;                                // slack rows read with zero weights
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- pass A: warp = residue n2, lane = frame ----
    {
        const int n2 = warp;
        cpx a[25];
        const float2 *ap = reinterpret_cast<const float2 *>(s_audio + lane * (HOP + 2) + 2 * n2);
#pragma unroll
        for (int n1 = 0; n1 < 25; ++n1) {
            const float2 x = ap[(16 * n1 + 2 * ((16 * n1) / HOP)) / 2];      // compile-time offset
            const float2 w = c_win2[8 * n1 + n2];
            a[n1] = {x.x * w.x, x.y * w.y};
        }
        dft25(a, c_tw25);
        if (n2 != 0) {
#pragma unroll
            for (int k1 = 1; k1 < 25; ++k1) a[k1] = cmul(a[k1], c_tw200[n2 * 25 + k1]);
        }
#pragma unroll
        for (int k1 = 0; k1 < 25; ++k1) s_Y[(n2 * 25 + k1) * 32 + lane] = make_float2(a[k1].re, a[k1].im);
    }
    __syncthreads();

    // ---- pass B: 13 items (k1 = j and 25 - j) per frame: radix-8, split, power (x4) ----
    for (int j = warp; j < 13; j += kWarps) {
        cpx y[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
            const float2 v = s_Y[(n2 * 25 + j) * 32 + lane];
            y[n2] = {v.x, v.y};
        }
        dft8(y);
        if (j == 0) {
            float pk, pm;
            split_power_x4(y[0], y[0], c_tw400[0], pk, pm);
            s_P[0 * 32 + lane] = pk; s_P[200 * 32 + lane] = pm;
#pragma unroll
            for (int k2 = 1; k2 < 4; ++k2) {
                split_power_x4(y[k2], y[8 - k2], c_tw400[25 * k2], pk, pm);
                s_P[(25 * k2) * 32 + lane] = pk; s_P[(200 - 25 * k2) * 32 + lane] = pm;
            }
            split_power_x4(y[4], y[4], c_tw400[100], pk, pm);
            s_P[100 * 32 + lane] = pk;
        } else {
            cpx y2[8];
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) {
                const float2 v = s_Y[(n2 * 25 + (25 - j)) * 32 + lane];
                y2[n2] = {v.x, v.y};
            }
            dft8(y2);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const int k = j + 25 * k2;
                float pk, pm;
                split_power_x4(y[k2], y2[7 - k2], c_tw400[k], pk, pm);
                s_P[k * 32 + lane] = pk; s_P[(200 - k) * 32 + lane] = pm;
            }
        }
    }
    __syncthreads();

    // ---- mel bank (4 bins per iteration, float4 weights) + dB ----
    const bool frame_ok = (t0 + lane) < t_full;
    float vmax = -INFINITY;
    const int ostride = n_mels + 1;
    for (int m = warp; m < n_mels; m += kWarps) {
        const float *pp = s_P + s_lo[m] * 32 + lane;
        const float4 *wp = s_w4 + s_woff4[m];
        const int n4 = s_cnt4[m];
        float acc = 0.f;
        for (int i = 0; i < n4; ++i) {
            const float4 w = wp[i];
            acc = fmaf(pp[0], w.x, acc);
            acc = fmaf(pp[32], w.y, acc);
            acc = fmaf(pp[64], w.z, acc);
            acc = fmaf(pp[96], w.w, acc);
            pp += 128;
        }
        const float db = kTenLog10Of2 * __log2f(pmax(0.25f * acc, kAmin));
        s_O[lane * ostride + m] = db;
        if (frame_ok) vmax = pmax(vmax, db);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = pmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    __shared__ float s_wmax[kWarps];
    if (lane == 0) s_wmax[warp] = vmax;
    __syncthreads();
    if (tid == 0) {
        float m = s_wmax[0];
#pragma unroll
        for (int i = 1; i < kWarps; ++i) m = pmax(m, s_wmax[i]);
        atomicMax(max_buf + b, f2ord(m));
    }
    const int nf = min(kTile, t_full - t0);
    float *dst = logmel + ((int64_t)b * t_full_max + t0) * n_mels;
    for (int f = warp; f < nf; f += kWarps)
        for (int m = lane; m < n_mels; m += 32) dst[f * n_mels + m] = s_O[f * ostride + m];
}

// ------------------------------------------------------------------ pass 2 --
// One CTA = (utterance b, chunk of kChunk frames); thread = mel column m.
constexpr int kChunk = 32;

template <typename T>
__global__ void __launch_bounds__(128)
finish_kernel(const float *__restrict__ logmel, const unsigned int *__restrict__ max_buf,
              const int32_t *__restrict__ wav_len, int64_t n_max, int hop, int n_mels, int deltas, int truncate,
              int t_full_max, T *__restrict__ out, int t_out, int32_t *__restrict__ out_frames) {
    const int b = blockIdx.y;
    const int m = threadIdx.x;
    const int64_t len_b = wav_len ? (int64_t)wav_len[b] : n_max;
    const int t_full = 1 + (int)(len_b / hop);
    int t_keep = t_full;
    if (truncate) t_keep = min(t_full, (int)((len_b + hop / 2) / hop));
    t_keep = min(t_keep, t_out);
    if (blockIdx.x == 0 && threadIdx.x == 0 && out_frames) out_frames[b] = t_keep;
    if (m >= n_mels) return;
    const int D = deltas ? 3 * n_mels : n_mels;
    const int tb = blockIdx.x * kChunk, te = min(tb + kChunk, t_out);
    const float floor_db = ord2f(max_buf[b]) - kTopDb;
    const float *x = logmel + (int64_t)b * t_full_max * n_mels + m;
    T *o = out + ((int64_t)b * t_out) * D + m;
    auto X = [&](int t) { return pmax(__ldg(x + (int64_t)min(max(t, 0), t_full - 1) * n_mels), floor_db); };
    // delta at clamped index tau, replicate padding of the input: sum_j j * x[clamp(tau + j)] / 10
    auto DELTA = [&](int tau) {
        tau = min(max(tau, 0), t_full - 1);
        return (-2.f * X(tau - 2) - X(tau - 1) + X(tau + 1) + 2.f * X(tau + 2)) / 10.f;
    };
    if (!deltas) {
        for (int t = tb; t < te; ++t) o[(int64_t)t * D] = from_f32<T>(t < t_keep ? X(t) : 0.f);
        return;
    }
    // sliding window of deltas d[k] = DELTA(t - 2 + k)
    float d0 = DELTA(tb - 2), d1 = DELTA(tb - 1), d2 = DELTA(tb), d3 = DELTA(tb + 1), d4;
    for (int t = tb; t < te; ++t) {
        d4 = DELTA(t + 2);
        const bool live = t < t_keep;
        const float dd = (-2.f * d0 - d1 + d3 + 2.f * d4) / 10.f;
        T *ot = o + (int64_t)t * D;
        ot[0] = from_f32<T>(live ? X(t) : 0.f);
        ot[n_mels] = from_f32<T>(live ? d2 : 0.f);
        ot[2 * n_mels] = from_f32<T>(live ? dd : 0.f);
        d0 = d1; d1 = d2; d2 = d3; d3 = d4;
    }
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

struct mlvae_fbank_plan {
    int sample_rate, hop, n_fft, n_mels, deltas;
    int skew_shift;
    int nnz;
    int *d_lo, *d_cnt, *d_woff;
    float *d_w;
    size_t smem_bytes;
    // fast path (hop in {160, 320}): weights padded to groups of 4
    int n4;
    int *d_cnt4, *d_woff4;
    float4 *d_w4;
    size_t smem_fast;
};

namespace {

void default_melmat(int sample_rate, int n_fft, int n_mels, std::vector<float> &fb) {
    // SpeechBrain Filterbank construction in float32 (python-float end points, float32 linspace/pow).
    const int n_stft = n_fft / 2 + 1;
    const double mel_lo = 2595.0 * std::log10(1.0 + 0.0 / 700.0);
    const double mel_hi = 2595.0 * std::log10(1.0 + (sample_rate / 2.0) / 700.0);
    std::vector<float> hz(n_mels + 2);
    const float step = (float)((mel_hi - mel_lo) / (n_mels + 1));
    for (int i = 0; i < n_mels + 2; ++i) {
        // torch.linspace: symmetric evaluation from both ends
        float mel = (i < (n_mels + 2) / 2) ? (float)mel_lo + step * (float)i : (float)mel_hi - step * (float)(n_mels + 1 - i);
        hz[i] = 700.f * (std::pow(10.f, mel / 2595.f) - 1.f);
    }
    fb.assign((size_t)n_stft * n_mels, 0.f);
    const float fstep = (float)((sample_rate / 2) / (double)(n_stft - 1));
    for (int k = 0; k < n_stft; ++k) {
        float f = (k < n_stft / 2) ? fstep * (float)k : (float)(sample_rate / 2) - fstep * (float)(n_stft - 1 - k);
        for (int m = 0; m < n_mels; ++m) {
            const float band = hz[m + 1] - hz[m];
            const float slope = (f - hz[m + 1]) / band;
            const float v = std::fmin(slope + 1.f, -slope + 1.f);
            fb[(size_t)k * n_mels + m] = v > 0.f ? v : 0.f;
        }
    }
}

}  // namespace

extern "C" {

int mlvae_fbank_plan_create(mlvae_fbank_plan **plan, int sample_rate, int hop_samples, int n_fft, int n_mels,
                            int deltas, const float *h_window, const float *h_melmat) {
    MLVAE_REQUIRE(plan, MLVAE_ERR_INVALID_ARG, "fbank_plan_create: plan is null");
    *plan = nullptr;
    MLVAE_REQUIRE(n_fft == kNfft, MLVAE_ERR_UNSUPPORTED,
                  "fbank: only n_fft == win_length == 400 (the reference geometry, run.yaml:26-29) is implemented, got %d", n_fft);
    MLVAE_REQUIRE(sample_rate > 0 && hop_samples > 0 && hop_samples <= 4096, MLVAE_ERR_INVALID_ARG, "fbank: bad sample_rate/hop");
    MLVAE_REQUIRE(n_mels > 0 && n_mels <= 128, MLVAE_ERR_UNSUPPORTED, "fbank: n_mels must be in [1, 128], got %d", n_mels);

    std::vector<float> win(kNfft);
    if (h_window) std::memcpy(win.data(), h_window, sizeof(float) * kNfft);
    else
        for (int n = 0; n < kNfft; ++n) win[n] = (float)(0.54 - 0.46 * std::cos(2.0 * M_PI * n / kNfft));
    std::vector<cpx> tw25(25), tw200(200), tw400(kBins);
    for (int q = 0; q < 5; ++q)
        for (int r = 0; r < 5; ++r) {
            const double a = -2.0 * M_PI * (q * r) / 25.0;
            tw25[q * 5 + r] = {(float)std::cos(a), (float)std::sin(a)};
        }
    for (int n2 = 0; n2 < 8; ++n2)
        for (int k1 = 0; k1 < 25; ++k1) {
            const double a = -2.0 * M_PI * (n2 * k1) / 200.0;
            tw200[n2 * 25 + k1] = {(float)std::cos(a), (float)std::sin(a)};
        }
    for (int k = 0; k < kBins; ++k) {
        const double a = -2.0 * M_PI * k / 400.0;
        tw400[k] = {(float)std::cos(a), (float)std::sin(a)};
    }
    MLVAE_CHECK_CUDA(cudaMemcpyToSymbol(c_win, win.data(), sizeof(float) * kNfft));
    MLVAE_CHECK_CUDA(cudaMemcpyToSymbol(c_win2, win.data(), sizeof(float) * kNfft));
    MLVAE_CHECK_CUDA(cudaMemcpyToSymbol(c_tw25, tw25.data(), sizeof(cpx) * 25));
    MLVAE_CHECK_CUDA(cudaMemcpyToSymbol(c_tw200, tw200.data(), sizeof(cpx) * 200));
    MLVAE_CHECK_CUDA(cudaMemcpyToSymbol(c_tw400, tw400.data(), sizeof(cpx) * kBins));

    std::vector<float> fb;
    if (h_melmat) fb.assign(h_melmat, h_melmat + (size_t)kBins * n_mels);
    else default_melmat(sample_rate, n_fft, n_mels, fb);
    std::vector<int> lo(n_mels), cnt(n_mels), woff(n_mels);
    std::vector<float> w;
    for (int m = 0; m < n_mels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < kBins; ++k)
            if (fb[(size_t)k * n_mels + m] != 0.f) { if (first < 0) first = k; last = k; }
        lo[m] = first < 0 ? 0 : first;
        cnt[m] = first < 0 ? 0 : last - first + 1;
        woff[m] = (int)w.size();
        for (int k = 0; k < cnt[m]; ++k) w.push_back(fb[(size_t)(lo[m] + k) * n_mels + m]);
    }
    if (w.empty()) w.push_back(0.f);

    auto *p = new mlvae_fbank_plan();
    p->sample_rate = sample_rate; p->hop = hop_samples; p->n_fft = n_fft; p->n_mels = n_mels; p->deltas = deltas ? 1 : 0;
    int s = 0;
    while (((hop_samples >> s) & 1) == 0) ++s;
    p->skew_shift = s;            // 0 => odd hop, already conflict free, no skew
    p->nnz = (int)w.size();
    const int span = (kTile - 1) * hop_samples + kNfft;
    const int span_sk = (s > 0 ? span + (span >> s) : span) + 1;
    p->smem_bytes = 200 * 32 * 8 + (size_t)kBins * 32 * 4 + (size_t)((span_sk + 3) & ~3) * 4 + (size_t)p->nnz * 4 + (size_t)n_mels * 12;
    if (p->smem_bytes > 227 * 1024) {
        const size_t need = p->smem_bytes;
        delete p;
        return fail(MLVAE_ERR_UNSUPPORTED, "fbank: hop/mel configuration needs %zu bytes of shared memory (> 227 KB)", need);
    }
    cudaError_t e = cudaSuccess;
    if ((e = cudaMalloc(&p->d_lo, sizeof(int) * n_mels)) != cudaSuccess ||
        (e = cudaMalloc(&p->d_cnt, sizeof(int) * n_mels)) != cudaSuccess ||
        (e = cudaMalloc(&p->d_woff, sizeof(int) * n_mels)) != cudaSuccess ||
        (e = cudaMalloc(&p->d_w, sizeof(float) * w.size())) != cudaSuccess) {
        delete p;
        return fail(MLVAE_ERR_CUDA, "fbank_plan_create: cudaMalloc: %s", cudaGetErrorString(e));
    }
    MLVAE_CHECK_CUDA(cudaMemcpy(p->d_lo, lo.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    MLVAE_CHECK_CUDA(cudaMemcpy(p->d_cnt, cnt.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    MLVAE_CHECK_CUDA(cudaMemcpy(p->d_woff, woff.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
    MLVAE_CHECK_CUDA(cudaMemcpy(p->d_w, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
    MLVAE_CHECK_CUDA(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
    // ---- fast-path tables ----
    p->n4 = 0; p->d_cnt4 = p->d_woff4 = nullptr; p->d_w4 = nullptr; p->smem_fast = 0;
    if (hop_samples == 160 || hop_samples == 320) {
        std::vector<int> cnt4(n_mels), woff4(n_mels);
        std::vector<float> w4;
        for (int m = 0; m < n_mels; ++m) {
            cnt4[m] = (cnt[m] + 3) / 4;
            woff4[m] = (int)w4.size() / 4;
            for (int k = 0; k < cnt4[m] * 4; ++k) w4.push_back(k < cnt[m] ? w[woff[m] + k] : 0.f);
        }
        bool fits = true;
        for (int m = 0; m < n_mels; ++m) fits = fits && (lo[m] + cnt4[m] * 4 <= kBins + 3);
        if (w4.empty()) { w4.assign(4, 0.f); }
        p->n4 = (int)w4.size() / 4;
        const size_t audio = hop_samples == 160 ? FastCfg<160>::kAudioFloats : FastCfg<320>::kAudioFloats;
        p->smem_fast = 200 * 32 * 8 + (size_t)(kBins + 3) * 32 * 4 + audio * 4 + (size_t)p->n4 * 16 + (size_t)n_mels * 12;
        if (fits && p->smem_fast <= 227 * 1024) {
            if ((e = cudaMalloc(&p->d_cnt4, sizeof(int) * n_mels)) != cudaSuccess ||
                (e = cudaMalloc(&p->d_woff4, sizeof(int) * n_mels)) != cudaSuccess ||
                (e = cudaMalloc(&p->d_w4, sizeof(float) * w4.size())) != cudaSuccess)
                return fail(MLVAE_ERR_CUDA, "fbank_plan_create: cudaMalloc: %s", cudaGetErrorString(e));
            MLVAE_CHECK_CUDA(cudaMemcpy(p->d_cnt4, cnt4.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
            MLVAE_CHECK_CUDA(cudaMemcpy(p->d_woff4, woff4.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
            MLVAE_CHECK_CUDA(cudaMemcpy(p->d_w4, w4.data(), sizeof(float) * w4.size(), cudaMemcpyHostToDevice));
            MLVAE_CHECK_CUDA(cudaFuncSetAttribute(logmel_fast_kernel<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            MLVAE_CHECK_CUDA(cudaFuncSetAttribute(logmel_fast_kernel<320>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        } else {
            p->n4 = 0;
        }
    }
    *plan = p;
    return MLVAE_OK;
}

int mlvae_fbank_plan_destroy(mlvae_fbank_plan *p) {
    if (!p) return MLVAE_OK;
    cudaFree(p->d_lo); cudaFree(p->d_cnt); cudaFree(p->d_woff); cudaFree(p->d_w);
    cudaFree(p->d_cnt4); cudaFree(p->d_woff4); cudaFree(p->d_w4);
    delete p;
    return MLVAE_OK;
}

int mlvae_fbank_frames(const mlvae_fbank_plan *p, int64_t n, int truncate_kaldi) {
    if (!p || n < 0) return MLVAE_ERR_INVALID_ARG;
    const int64_t full = 1 + n / p->hop;
    const int64_t kaldi = (n + p->hop / 2) / p->hop;
    return (int)((truncate_kaldi && kaldi < full) ? kaldi : full);
}

int mlvae_fbank_feature_dim(const mlvae_fbank_plan *p) {
    if (!p) return MLVAE_ERR_INVALID_ARG;
    return p->deltas ? 3 * p->n_mels : p->n_mels;
}

size_t mlvae_fbank_scratch_bytes(const mlvae_fbank_plan *p, int B, int64_t n_max) {
    if (!p || B <= 0 || n_max < 0) return 0;
    const size_t t_full_max = (size_t)(1 + n_max / p->hop);
    return 256 * ((sizeof(unsigned int) * (size_t)B + 255) / 256) + sizeof(float) * (size_t)B * t_full_max * p->n_mels;
}

int mlvae_fbank_fwd(const mlvae_fbank_plan *p, const float *d_wav, const int32_t *d_wav_len, int B, int64_t n_max,
                    int64_t n_stride, int truncate_kaldi, void *d_out, int out_dtype, int t_out,
                    int32_t *d_out_frames, void *d_scratch, void *stream) {
    MLVAE_REQUIRE(p && d_wav && d_out && d_scratch, MLVAE_ERR_INVALID_ARG, "fbank_fwd: missing buffers");
    MLVAE_REQUIRE(B > 0 && B <= 65535 && n_max >= 0 && n_stride >= n_max && t_out > 0, MLVAE_ERR_INVALID_ARG, "fbank_fwd: bad sizes");
    MLVAE_REQUIRE(n_max / p->hop < (1 << 24), MLVAE_ERR_UNSUPPORTED, "fbank_fwd: utterance too long");
    cudaStream_t st = (cudaStream_t)stream;
    const int t_full_max = 1 + (int)(n_max / p->hop);
    unsigned int *max_buf = (unsigned int *)d_scratch;
    float *logmel = (float *)((char *)d_scratch + 256 * ((sizeof(unsigned int) * (size_t)B + 255) / 256));
    MLVAE_CHECK_CUDA(cudaMemsetAsync(max_buf, 0, sizeof(unsigned int) * B, st));
    MelTables mel{p->d_lo, p->d_cnt, p->d_woff, p->d_w, p->nnz};
    const int vec4_ok = (p->hop % 4 == 0) && (n_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_wav) & 15u) == 0);
    dim3 g1((t_full_max + kTile - 1) / kTile, B);
    const bool fast = p->n4 > 0 && (n_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(d_wav) & 7u) == 0);
    if (fast) {
        MelTables4 m4{p->d_lo, p->d_cnt4, p->d_woff4, p->d_w4, p->n4};
        if (p->hop == 160)
            logmel_fast_kernel<160><<<g1, kThreadsFb, p->smem_fast, st>>>(d_wav, d_wav_len, n_max, n_stride, p->n_mels, m4, logmel, t_full_max, max_buf);
        else
            logmel_fast_kernel<320><<<g1, kThreadsFb, p->smem_fast, st>>>(d_wav, d_wav_len, n_max, n_stride, p->n_mels, m4, logmel, t_full_max, max_buf);
    } else {
        logmel_kernel<<<g1, kThreadsFb, p->smem_bytes, st>>>(d_wav, d_wav_len, n_max, n_stride, p->hop, p->skew_shift,
                                                             p->n_mels, mel, logmel, t_full_max, max_buf, vec4_ok);
    }
    MLVAE_CHECK_CUDA(cudaGetLastError());
    dim3 g2((t_out + kChunk - 1) / kChunk, B);
    if (out_dtype == MLVAE_F32)
        finish_kernel<float><<<g2, 128, 0, st>>>(logmel, max_buf, d_wav_len, n_max, p->hop, p->n_mels, p->deltas,
                                                 truncate_kaldi, t_full_max, (float *)d_out, t_out, d_out_frames);
    else if (out_dtype == MLVAE_BF16)
        finish_kernel<__nv_bfloat16><<<g2, 128, 0, st>>>(logmel, max_buf, d_wav_len, n_max, p->hop, p->n_mels, p->deltas,
                                                         truncate_kaldi, t_full_max, (__nv_bfloat16 *)d_out, t_out, d_out_frames);
    else return fail(MLVAE_ERR_INVALID_ARG, "unknown dtype %d", out_dtype);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

}  // extern "C"
