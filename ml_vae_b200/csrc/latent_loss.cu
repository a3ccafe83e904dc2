// Fused latent-block loss kernels for sm_100a:
//   reparameterise + KL (+ masked mean)   fwd / bwd     vanilla_vae.py:37-45 + data_utils.py:67-104
//   Gaussian-NLL / MSE reconstruction     fwd / bwd     decoder.py:37-53     + data_utils.py:67-104
//   stand-alone length-masked reduction   fwd / bwd     data_utils.py:67-104
//   Philox eps materialisation                              (replaces torch.randn_like, vanilla_vae.py:39)
//
// All of them are HBM-streaming kernels: 16-byte vector loads/stores, one pass
// over the data, warp-shuffle -> CTA -> per-CTA partial -> fixed-order final sum
// by the last CTA (deterministic, no float atomics).  Roofline: HBM; algorithmic
// bytes per frame are in DESIGN.md section "Kernels".
#include "common.cuh"
#include "philox.cuh"

namespace mlvae {
namespace {

constexpr int kThreads = 256;
constexpr float kLog2Pi_f32 = 1.8378770351409912f;   // float32(log(float32(2*pi))), decoder.py:42
constexpr float kReconEps = 1e-5f;                    // decoder.py:41

template <typename T> __device__ __forceinline__ float fast_exp(float x);
template <> __device__ __forceinline__ float fast_exp<float>(float x) { return expf(x); }
// bf16 kernels: one multiply + MUFU.EX2, flush-to-zero form (no denormal fix-up code around the MUFU: exp underflows to 0 below
// e^-87 instead of producing denormals, far below anything bf16 can carry)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <> __device__ __forceinline__ float fast_exp<__nv_bfloat16>(float x) { return ex2_ftz(x * 1.4426950408889634f); }
// exp(x / 2)
// 1 / x for x >= 1e-5: the bf16 kernels take MUFU.RCP (1 ulp), the float32 kernels the IEEE division
template <typename T> __device__ __forceinline__ float fast_rcp(float x);
template <> __device__ __forceinline__ float fast_rcp<float>(float x) { return 1.f / x; }
template <> __device__ __forceinline__ float fast_rcp<__nv_bfloat16>(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <typename T> __device__ __forceinline__ float exp_half(float x);
template <> __device__ __forceinline__ float exp_half<float>(float x) { return expf(0.5f * x); }
template <> __device__ __forceinline__ float exp_half<__nv_bfloat16>(float x) { return ex2_ftz(x * 0.7213475204444817f); }

// ---------------------------------------------------------------------------
// Walks the (rows x C) matrix in units of VEC contiguous elements of one row.
// Every thread keeps (row, column-vector, b, t) incrementally: no division in
// the loop.  rows = B*T, row = b*T + t.
// ---------------------------------------------------------------------------
struct RowWalker {
    int64_t vec, nvec;      // current / total vector index
    int64_t step;           // threads in the grid
    int vpr;                // vectors per row
    int cv;                 // column vector of the current item
    int b, t, T;
    int step_rows, step_cv; // step / vpr, step % vpr
    int srow_b, srow_t;     // step_rows / T, step_rows % T

    __device__ __forceinline__ RowWalker(int64_t rows, int vpr_, int T_) {
        vpr = vpr_; T = T_;
        nvec = rows * vpr;
        step = (int64_t)gridDim.x * blockDim.x;
        vec = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const int64_t row = vec / vpr;
        cv = (int)(vec - row * vpr);
        b = (int)(row / T);
        t = (int)(row - (int64_t)b * T);
        step_rows = (int)(step / vpr);
        step_cv = (int)(step - (int64_t)step_rows * vpr);
        srow_b = step_rows / T;
        srow_t = step_rows - srow_b * T;
    }
    __device__ __forceinline__ bool valid() const { return vec < nvec; }
    __device__ __forceinline__ void next() {
        vec += step;
        cv += step_cv;
        int carry = 0;
        if (cv >= vpr) { cv -= vpr; carry = 1; }
        b += srow_b;
        t += srow_t + carry;
        if (t >= T) { t -= T; ++b; }
        if (t >= T) { t -= T; ++b; }
    }
};

// Final, deterministic reduction: the last CTA to arrive sums the per-CTA
// partials in index order (double accumulation) and writes {mean, sum, count}.
__device__ __forceinline__ bool last_cta_arrives(ReduceScratch *s) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&s->ticket, 1u);
        s_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    return s_last;
}

__device__ __forceinline__ double ordered_sum(const float *p, int n) {
    // warp 0 only: lane-strided double sums, then a fixed butterfly
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) acc += (double)__ldcg(p + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

__device__ __forceinline__ int64_t total_valid_frames(const float *lens, int B, int T) {
    // warp 0 only
    int64_t n = 0;
    for (int b = threadIdx.x; b < B; b += 32) n += valid_frames(__ldg(lens + b), T);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    return n;
}

__device__ __forceinline__ void finish_masked_mean(float local, ReduceScratch *s, const float *lens, int B,
                                                   int T, int C, float *out) {
    const float bs = block_sum(local);
    if (threadIdx.x == 0) s->partial[blockIdx.x] = bs;
    if (last_cta_arrives(s)) {
        if (threadIdx.x < 32) {
            const double tot = ordered_sum(s->partial, gridDim.x);
            const int64_t frames = total_valid_frames(lens, B, T);
            if (threadIdx.x == 0) {
                const float cnt = (float)(frames * (int64_t)C);
                out[0] = (float)tot / cnt;       // 0/0 -> NaN exactly like sum/sum(mask) in the reference
                out[1] = (float)tot;
                out[2] = cnt;
                s->ticket = 0;                   // leave the scratch ready for the next launch
            }
        }
    }
}

// 1 / (count*C) * upstream scalar, or 0 when no reduced-loss gradient flows.
__device__ __forceinline__ float mean_grad_scale(const float *g_mean, const float *lens, int B, int T, int C) {
    __shared__ float s_scale;
    if (g_mean == nullptr) return 0.f;
    if (threadIdx.x < 32) {
        const int64_t frames = total_valid_frames(lens, B, T);
        if (threadIdx.x == 0) s_scale = __ldg(g_mean) / (float)(frames * (int64_t)C);
    }
    __syncthreads();
    return s_scale;
}

// ------------------------------------------------------------- Philox ------
__global__ void __launch_bounds__(kThreads) philox_u32_kernel(uint64_t seed, uint64_t offset, int64_t n, uint32_t *out) {
    const PhiloxKey key(seed);
    const int64_t nblk = (n + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nblk; q += (int64_t)gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)((uint64_t)q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)), key);
        const uint32_t v[4] = {r.x, r.y, r.z, r.w};
        for (int k = 0; k < 4; ++k)
            if (4 * q + k < n) out[4 * q + k] = v[k];
    }
}

// kV2: the eps stream of the bf16 kernels (8 normals per Philox call, philox.cuh) instead of the float32 kernels' stream
template <typename T, bool kV2>
__global__ void __launch_bounds__(kThreads) philox_normal_kernel(uint64_t seed, uint64_t offset, int64_t n, T *out) {
    const PhiloxKey key(seed);
    constexpr int PB = kV2 ? 8 : 4;
    const int64_t nblk = (n + PB - 1) / PB;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nblk; q += (int64_t)gridDim.x * blockDim.x) {
        float v[8];
        if constexpr (kV2) philox_normal8((uint64_t)q, offset, key, v);
        else philox_normal4((uint64_t)q, offset, key, v);
        for (int k = 0; k < PB; ++k)
            if (PB * q + k < n) out[PB * q + k] = from_f32<T>(v[k]);
    }
}

// eps for VEC (= 1, 4 or 8) consecutive stream elements starting at element e0: the float32 kernels draw the v1 stream
// (4 normals per Philox call), the bf16 kernels the v2 stream (8 per call).
template <typename T, int VEC>
__device__ __forceinline__ void stream_eps(int64_t e0, uint64_t offset, const PhiloxKey &key, float *eps) {
    if constexpr (sizeof(T) == 2) {
        if constexpr (VEC == 8) {
            philox_normal8((uint64_t)e0 >> 3, offset, key, eps);
        } else {
            static_assert(VEC == 1, "bf16 kernels are instantiated with VEC = 8 or 1");
            float v[8];
            philox_normal8((uint64_t)e0 >> 3, offset, key, v);
            eps[0] = v[e0 & 7];
        }
    } else if constexpr (VEC == 1) {
        float v[4];
        philox_normal4((uint64_t)e0 >> 2, offset, key, v);
        eps[0] = v[e0 & 3];
    } else {
#pragma unroll
        for (int j = 0; j < VEC / 4; ++j) philox_normal4(((uint64_t)e0 >> 2) + j, offset, key, eps + 4 * j);
    }
}

template <typename T, int VEC> struct Chunk {   // VEC == Vec<T>::N (vector path) or 1 (scalar path)
    float v[VEC];
    __device__ __forceinline__ void load(const T *p) {
        if constexpr (VEC == 1) v[0] = to_f32<T>(__ldg(p));
        else { Vec<T> x; x.load_stream(p);
#pragma unroll
               for (int i = 0; i < VEC; ++i) v[i] = x.v[i]; }
    }
    __device__ __forceinline__ void store(T *p) const {
        if constexpr (VEC == 1) p[0] = from_f32<T>(v[0]);
        else { Vec<T> x;
#pragma unroll
               for (int i = 0; i < VEC; ++i) x.v[i] = v[i];
               x.store(p); }
    }
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = 0.f;
    }
};

// ------------------------------------------------ reparameterise + KL ------
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
reparam_kl_fwd_kernel(const T *__restrict__ mu, const T *__restrict__ logvar, const T *__restrict__ eps,
                      const PhiloxKey key, uint64_t offset, const uint64_t *__restrict__ offset_add, const float *__restrict__ lens, int B, int Tn, int L,
                      int64_t ld_in, T *__restrict__ z, T *__restrict__ kl_elem, float *__restrict__ kl_out, ReduceScratch *scratch) {
    if (offset_add) offset += __ldg(offset_add);      // device-resident step counter (CUDA-graph friendly)
    float acc = 0.f;
    for (RowWalker w((int64_t)B * Tn, L / VEC, Tn); w.valid(); w.next()) {
        const int64_t e0 = w.vec * VEC;
        const int64_t ei = ((int64_t)w.b * Tn + w.t) * ld_in + (int64_t)w.cv * VEC;      // mu / logvar rows are ld_in elements apart
        Chunk<T, VEC> m, lv, ep, zz, kk;
        m.load(mu + ei);
        lv.load(logvar + ei);
        if (eps) ep.load(eps + e0);
        else stream_eps<T, VEC>(e0, offset, key, ep.v);
        float maskf = 1.f;
        if (kl_out) maskf = ((float)w.t < mask_threshold(__ldg(lens + w.b), Tn)) ? 1.f : 0.f;
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float sd = exp_half<T>(lv.v[i]);
            zz.v[i] = fmaf(ep.v[i], sd, m.v[i]);
            float k;
            if constexpr (sizeof(T) == 2) k = fmaf(0.5f, fmaf(m.v[i], m.v[i], sd * sd), fmaf(-0.5f, lv.v[i], -0.5f));    // same value, 4 instructions
            else k = -0.5f * (1.f + lv.v[i] - m.v[i] * m.v[i] - sd * sd);
            kk.v[i] = k;
            part = fmaf(k, maskf, part);      // multiply, not select: inf*0 = NaN like loss*mask in the reference
        }
        acc += part;
        zz.store(z + e0);
        if (kl_elem) kk.store(kl_elem + e0);
    }
    if (kl_out) finish_masked_mean(acc, scratch, lens, B, Tn, L, kl_out);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
reparam_kl_bwd_kernel(const T *__restrict__ mu, const T *__restrict__ logvar, const T *__restrict__ eps,
                      const PhiloxKey key, uint64_t offset, const uint64_t *__restrict__ offset_add, const T *__restrict__ grad_z,
                      const T *__restrict__ grad_kl_elem, const float *__restrict__ grad_kl_mean,
                      const float *__restrict__ lens, int B, int Tn, int L, int64_t ld_in, int64_t ld_out,
                      T *__restrict__ grad_mu, T *__restrict__ grad_logvar) {
    if (offset_add) offset += __ldg(offset_add);
    const float gscale = mean_grad_scale(grad_kl_mean, lens, B, Tn, L);
    for (RowWalker w((int64_t)B * Tn, L / VEC, Tn); w.valid(); w.next()) {
        const int64_t e0 = w.vec * VEC;
        const int64_t row = (int64_t)w.b * Tn + w.t;
        const int64_t ei = row * ld_in + (int64_t)w.cv * VEC, eo = row * ld_out + (int64_t)w.cv * VEC;
        Chunk<T, VEC> m, lv, ep, gz, ge, gm, gl;
        m.load(mu + ei);
        lv.load(logvar + ei);
        if (grad_z) gz.load(grad_z + e0); else gz.zero();
        if (grad_kl_elem) ge.load(grad_kl_elem + e0); else ge.zero();
        if (eps) ep.load(eps + e0);
        else stream_eps<T, VEC>(e0, offset, key, ep.v);
        float gm_row = 0.f;
        if (grad_kl_mean) gm_row = gscale * (((float)w.t < mask_threshold(__ldg(lens + w.b), Tn)) ? 1.f : 0.f);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float sd = exp_half<T>(lv.v[i]);
            const float g = ge.v[i] + gm_row;
            gm.v[i] = fmaf(g, m.v[i], gz.v[i]);
            gl.v[i] = fmaf(gz.v[i] * 0.5f * sd, ep.v[i], g * 0.5f * (sd * sd - 1.f));
        }
        gm.store(grad_mu + eo);
        gl.store(grad_logvar + eo);
    }
}

// ------------------------------------------------ reconstruction loss ------
template <typename T, int VEC, bool kMse>
__global__ void __launch_bounds__(kThreads)
recon_fwd_kernel(const T *__restrict__ mean, const T *__restrict__ logvar, const T *__restrict__ target,
                 const float *__restrict__ lens, int B, int Tn, int D, T *__restrict__ elem,
                 float *__restrict__ out, ReduceScratch *scratch) {
    float acc = 0.f;
    for (RowWalker w((int64_t)B * Tn, D / VEC, Tn); w.valid(); w.next()) {
        const int64_t e0 = w.vec * VEC;
        Chunk<T, VEC> m, lv, tg, ll;
        m.load(mean + e0);
        tg.load(target + e0);
        if constexpr (!kMse) lv.load(logvar + e0);
        float maskf = 1.f;
        if (out) maskf = ((float)w.t < mask_threshold(__ldg(lens + w.b), Tn)) ? 1.f : 0.f;
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float d = tg.v[i] - m.v[i];
            float l;
            if constexpr (kMse) l = d * d;
            else if constexpr (sizeof(T) == 2) l = 0.5f * (kLog2Pi_f32 + lv.v[i] + d * d * fast_rcp<T>(fast_exp<T>(lv.v[i]) + kReconEps));
            else l = 0.5f * (kLog2Pi_f32 + lv.v[i] + d * d / (fast_exp<T>(lv.v[i]) + kReconEps));
            ll.v[i] = l;
            part = fmaf(l, maskf, part);
        }
        acc += part;
        if (elem) ll.store(elem + e0);
    }
    if (out) finish_masked_mean(acc, scratch, lens, B, Tn, D, out);
}

template <typename T, int VEC, bool kMse>
__global__ void __launch_bounds__(kThreads)
recon_bwd_kernel(const T *__restrict__ mean, const T *__restrict__ logvar, const T *__restrict__ target,
                 const T *__restrict__ grad_elem, const float *__restrict__ grad_mean_scalar,
                 const float *__restrict__ lens, int B, int Tn, int D,
                 T *__restrict__ grad_mean, T *__restrict__ grad_logvar, T *__restrict__ grad_target) {
    const float gscale = mean_grad_scale(grad_mean_scalar, lens, B, Tn, D);
    for (RowWalker w((int64_t)B * Tn, D / VEC, Tn); w.valid(); w.next()) {
        const int64_t e0 = w.vec * VEC;
        Chunk<T, VEC> m, lv, tg, ge, gm, gl, gt;
        m.load(mean + e0);
        tg.load(target + e0);
        if constexpr (!kMse) lv.load(logvar + e0);
        if (grad_elem) ge.load(grad_elem + e0); else ge.zero();
        float gm_row = 0.f;
        if (grad_mean_scalar) gm_row = gscale * (((float)w.t < mask_threshold(__ldg(lens + w.b), Tn)) ? 1.f : 0.f);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float g = ge.v[i] + gm_row;
            const float d = tg.v[i] - m.v[i];
            if constexpr (kMse) {
                gt.v[i] = 2.f * d * g;
                gm.v[i] = -gt.v[i];
            } else {
                const float e = fast_exp<T>(lv.v[i]);
                const float inv = fast_rcp<T>(e + kReconEps);
                gt.v[i] = g * d * inv;                       // d/dtarget = (t-m)/(e+eps)
                gm.v[i] = -gt.v[i];
                gl.v[i] = g * 0.5f * (1.f - d * d * e * inv * inv);
            }
        }
        gm.store(grad_mean + e0);
        if constexpr (!kMse) gl.store(grad_logvar + e0);
        if (grad_target) gt.store(grad_target + e0);
    }
}

// ------------------------------------------ stand-alone masked reduction ----
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
masked_sum_kernel(const T *__restrict__ loss, const float *__restrict__ lens, int B, int Tn, int C, int reduction,
                  float *__restrict__ out, ReduceScratch *scratch) {
    float acc = 0.f;
    for (RowWalker w((int64_t)B * Tn, C / VEC, Tn); w.valid(); w.next()) {
        Chunk<T, VEC> x;
        x.load(loss + w.vec * VEC);
        const float maskf = ((float)w.t < mask_threshold(__ldg(lens + w.b), Tn)) ? 1.f : 0.f;
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) part += x.v[i] * maskf;
        acc += part;
    }
    const float bs = block_sum(acc);
    if (threadIdx.x == 0) scratch->partial[blockIdx.x] = bs;
    if (last_cta_arrives(scratch) && threadIdx.x < 32) {
        const double tot = ordered_sum(scratch->partial, gridDim.x);
        const int64_t frames = total_valid_frames(lens, B, Tn);
        if (threadIdx.x == 0) {
            const float denom = (reduction == MLVAE_RED_MEAN) ? (float)(frames * (int64_t)C) : (float)B;
            out[0] = (float)tot / denom;
            scratch->ticket = 0;
        }
    }
}

// 'batch' reduction: one CTA per utterance row (B is small; T*C per row is what is summed).
template <typename T>
__global__ void __launch_bounds__(kThreads)
masked_rowmean_kernel(const T *__restrict__ loss, const float *__restrict__ lens, int Tn, int C, float *__restrict__ out) {
    const int b = blockIdx.x;
    const float thr = mask_threshold(__ldg(lens + b), Tn);
    const T *row = loss + (int64_t)b * Tn * C;
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < (int64_t)Tn * C; i += blockDim.x) {
        const int t = (int)(i / C);
        acc += to_f32<T>(__ldg(row + i)) * (((float)t < thr) ? 1.f : 0.f);
    }
    const float s = block_sum(acc);
    if (threadIdx.x == 0) out[b] = s / (float)((int64_t)valid_frames(__ldg(lens + b), Tn) * C);
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
masked_reduce_bwd_kernel(const float *__restrict__ grad_out, const float *__restrict__ lens, int B, int Tn, int C,
                         int reduction, T *__restrict__ grad_loss) {
    __shared__ float s_scale;
    if (reduction != MLVAE_RED_BATCH) {
        if (threadIdx.x < 32) {
            const int64_t frames = total_valid_frames(lens, B, Tn);
            if (threadIdx.x == 0)
                s_scale = __ldg(grad_out) / ((reduction == MLVAE_RED_MEAN) ? (float)(frames * (int64_t)C) : (float)B);
        }
        __syncthreads();
    }
    for (RowWalker w((int64_t)B * Tn, C / VEC, Tn); w.valid(); w.next()) {
        const float len_b = __ldg(lens + w.b);
        float g = ((float)w.t < mask_threshold(len_b, Tn)) ? 1.f : 0.f;
        if (reduction == MLVAE_RED_BATCH) g *= __ldg(grad_out + w.b) / (float)((int64_t)valid_frames(len_b, Tn) * C);
        else g *= s_scale;
        Chunk<T, VEC> x;
#pragma unroll
        for (int i = 0; i < VEC; ++i) x.v[i] = g;
        x.store(grad_loss + w.vec * VEC);
    }
}

// ------------------------------------ GMM-VAE reparameterise + KL vs a learned prior ------
// modules/gmm_vae.py:51-67: z = mu + exp(0.5 lv) eps ;
// kl = -0.5 (1 + lv - plv - (exp(lv) + (mu - pmu)^2) / (exp(plv) + 1e-5)), unreduced (the mixing by
// gmm_weight / pi that follows needs it per element).  Flat elementwise, 16-byte vectors.
constexpr float kGmmEps = 1e-5f;

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
gmm_reparam_kl_fwd_kernel(const T *__restrict__ mu, const T *__restrict__ logvar, const T *__restrict__ pmu,
                          const T *__restrict__ plogvar, const T *__restrict__ eps, uint64_t seed, uint64_t offset,
                          const uint64_t *__restrict__ offset_add, int64_t n, T *__restrict__ z, T *__restrict__ kl) {
    const PhiloxKey key(seed);
    if (offset_add) offset += __ldg(offset_add);
    const int64_t nvec = n / VEC;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e0 = v * VEC;
        Chunk<T, VEC> m, lv, pm, plv, ep, zz, kk;
        m.load(mu + e0); lv.load(logvar + e0); pm.load(pmu + e0); plv.load(plogvar + e0);
        if (eps) ep.load(eps + e0);
        else stream_eps<T, VEC>(e0, offset, key, ep.v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float sd = exp_half<T>(lv.v[i]);
            zz.v[i] = fmaf(ep.v[i], sd, m.v[i]);
            const float d = m.v[i] - pm.v[i];
            kk.v[i] = -0.5f * (1.f + lv.v[i] - plv.v[i] - (sd * sd + d * d) / (fast_exp<T>(plv.v[i]) + kGmmEps));
        }
        zz.store(z + e0);
        kk.store(kl + e0);
    }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
gmm_reparam_kl_bwd_kernel(const T *__restrict__ mu, const T *__restrict__ logvar, const T *__restrict__ pmu,
                          const T *__restrict__ plogvar, const T *__restrict__ eps, uint64_t seed, uint64_t offset,
                          const uint64_t *__restrict__ offset_add, const T *__restrict__ grad_z, const T *__restrict__ grad_kl,
                          int64_t n, T *__restrict__ g_mu, T *__restrict__ g_lv, T *__restrict__ g_pmu, T *__restrict__ g_plv) {
    const PhiloxKey key(seed);
    if (offset_add) offset += __ldg(offset_add);
    const int64_t nvec = n / VEC;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e0 = v * VEC;
        Chunk<T, VEC> m, lv, pm, plv, ep, gz, gk, om, ol, opm, opl;
        m.load(mu + e0); lv.load(logvar + e0); pm.load(pmu + e0); plv.load(plogvar + e0);
        if (grad_z) gz.load(grad_z + e0); else gz.zero();
        if (grad_kl) gk.load(grad_kl + e0); else gk.zero();
        if (eps) ep.load(eps + e0);
        else stream_eps<T, VEC>(e0, offset, key, ep.v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float sd = exp_half<T>(lv.v[i]);
            const float e = sd * sd, ep_ = fast_exp<T>(plv.v[i]);
            const float inv = 1.f / (ep_ + kGmmEps);
            const float d = m.v[i] - pm.v[i];
            const float g = gk.v[i];
            om.v[i] = fmaf(g, d * inv, gz.v[i]);
            opm.v[i] = -g * d * inv;
            ol.v[i] = fmaf(gz.v[i] * 0.5f * sd, ep.v[i], -0.5f * g * (1.f - e * inv));
            opl.v[i] = 0.5f * g * (1.f - (e + d * d) * ep_ * inv * inv);
        }
        om.store(g_mu + e0); ol.store(g_lv + e0); opm.store(g_pmu + e0); opl.store(g_plv + e0);
    }
}

// --------------------------------------------------- apply_weight (utils/data_utils.py:32-64) ------
// out[m, c] = sum_n w[m, n] * x[m, n, c]; the reference runs M tiny (1 x N) @ (N x C) bmm problems.
// One warp per row m, lanes over c.
template <typename T>
__global__ void __launch_bounds__(kThreads)
apply_weight_fwd_kernel(const T *__restrict__ x, const T *__restrict__ w, int64_t M, int N, int C, T *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M; m += warps) {
        const T *xm = x + m * N * C;
        for (int c = lane; c < C; c += 32) {
            float acc = 0.f;
            for (int nn = 0; nn < N; ++nn) acc = fmaf(to_f32<T>(w[m * N + nn]), to_f32<T>(xm[(int64_t)nn * C + c]), acc);
            out[m * C + c] = from_f32<T>(acc);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
apply_weight_bwd_kernel(const T *__restrict__ x, const T *__restrict__ w, const T *__restrict__ g, int64_t M, int N, int C,
                        T *__restrict__ gx, T *__restrict__ gw) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < M; m += warps) {
        const T *xm = x + m * N * C;
        for (int nn = 0; nn < N; ++nn) {
            const float wn = to_f32<T>(w[m * N + nn]);
            float dot = 0.f;
            for (int c = lane; c < C; c += 32) {
                const float gc = to_f32<T>(g[m * C + c]);
                if (gx) gx[(m * N + nn) * C + c] = from_f32<T>(wn * gc);
                dot = fmaf(to_f32<T>(xm[(int64_t)nn * C + c]), gc, dot);
            }
            dot = warp_sum(dot);
            if (gw && lane == 0) gw[m * N + nn] = from_f32<T>(dot);
        }
    }
}

// Vector path (C % VEC == 0, 16-byte aligned): thread = (row m, VEC consecutive columns); a row's C / VEC threads are
// consecutive lanes, so all N + 1 streams are read / written with coalesced 16-byte accesses and every thread keeps
// N independent loads in flight (the warp-per-row form above moves 2 bytes per lane and reaches 20 % of the HBM peak).
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
apply_weight_fwd_vec_kernel(const T *__restrict__ x, const T *__restrict__ w, int64_t M, int N, int C, T *__restrict__ out) {
    const int VC = C / VEC;
    const int64_t total = M * VC;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int64_t m = i / VC;
        const int vc = (int)(i - m * VC);
        Chunk<T, VEC> acc, xv;
        acc.zero();
        for (int nn = 0; nn < N; ++nn) {
            xv.load(x + (m * N + nn) * C + vc * VEC);
            const float wn = to_f32<T>(__ldg(w + m * N + nn));
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc.v[e] = fmaf(wn, xv.v[e], acc.v[e]);
        }
        acc.store(out + m * C + vc * VEC);
    }
}

// VC = C / VEC must be a power of two <= 32 here: the dot products of a row are reduced with shuffles inside its VC lanes.
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
apply_weight_bwd_vec_kernel(const T *__restrict__ x, const T *__restrict__ w, const T *__restrict__ g, int64_t M, int N, int C,
                            T *__restrict__ gx, T *__restrict__ gw) {
    const int VC = C / VEC;
    const int64_t total = M * VC;
    const int64_t span = (int64_t)gridDim.x * kThreads;
    const int64_t rounds = (total + span - 1) / span;          // every lane runs every round: the shuffles need full warps
    for (int64_t r = 0; r < rounds; ++r) {
        const int64_t i = r * span + (int64_t)blockIdx.x * kThreads + threadIdx.x;
        const bool ok = i < total;
        const int64_t m = ok ? i / VC : 0;
        const int vc = ok ? (int)(i - m * VC) : 0;
        Chunk<T, VEC> gv, xv, o;
        gv.zero();
        if (ok) gv.load(g + m * C + vc * VEC);
        for (int nn = 0; nn < N; ++nn) {
            float dot = 0.f;
            if (ok) {
                xv.load(x + (m * N + nn) * C + vc * VEC);
                const float wn = to_f32<T>(__ldg(w + m * N + nn));
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    o.v[e] = wn * gv.v[e];
                    dot = fmaf(xv.v[e], gv.v[e], dot);
                }
                if (gx) o.store(gx + (m * N + nn) * C + vc * VEC);
            }
            for (int s = VC >> 1; s > 0; s >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, s);
            if (ok && gw && vc == 0) gw[m * N + nn] = from_f32<T>(dot);
        }
    }
}

// ------------------------------------------------------------ launching ----
inline int grid_for(int64_t nvec, int ctas_per_sm = 8) {
    int64_t g = (nvec + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (g > cap) g = cap;
    if (g > kMaxPartials) g = kMaxPartials;
    if (g < 1) g = 1;
    return (int)g;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_btc(int B, int T, int C) {
    MLVAE_REQUIRE(B > 0 && T > 0 && C > 0, MLVAE_ERR_INVALID_ARG, "B, T, C must be positive (got %d, %d, %d)", B, T, C);
    MLVAE_REQUIRE((int64_t)B * T < (1LL << 31), MLVAE_ERR_UNSUPPORTED, "B*T must be < 2^31");
    return MLVAE_OK;
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

// Dispatch helper: vector path when the row length and every pointer allow 16-byte access.
#define MLVAE_DISPATCH(DT, C, ALIGNED, ...)                                                    \
    do {                                                                                       \
        if ((DT) == MLVAE_F32) {                                                               \
            using T = float;                                                                   \
            if ((ALIGNED) && (C) % 4 == 0) { constexpr int VEC = 4; __VA_ARGS__; }                    \
            else { constexpr int VEC = 1; __VA_ARGS__; }                                              \
        } else if ((DT) == MLVAE_BF16) {                                                       \
            using T = __nv_bfloat16;                                                           \
            if ((ALIGNED) && (C) % 8 == 0) { constexpr int VEC = 8; __VA_ARGS__; }                    \
            else { constexpr int VEC = 1; __VA_ARGS__; }                                              \
        } else                                                                                 \
            return fail(MLVAE_ERR_INVALID_ARG, "unknown dtype %d", (int)(DT));                 \
    } while (0)

extern "C" {

int mlvae_abi_version(void) { return MLVAE_ABI_VERSION; }
const char *mlvae_last_error(void) { return err_buf(); }

int mlvae_device_info(int *sms, int *cc) {
    int dev = 0, n = 0, maj = 0, min = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(MLVAE_ERR_NO_DEVICE, "no CUDA device");
    MLVAE_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    MLVAE_CHECK_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    MLVAE_CHECK_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sms) *sms = n;
    if (cc) *cc = maj * 10 + min;
    return MLVAE_OK;
}

size_t mlvae_reduce_scratch_bytes(void) { return sizeof(ReduceScratch); }

int mlvae_philox_u32(uint64_t seed, uint64_t offset, int64_t n, uint32_t *d_out, void *stream) {
    MLVAE_REQUIRE(d_out && n >= 0, MLVAE_ERR_INVALID_ARG, "philox_u32: bad arguments");
    if (n == 0) return MLVAE_OK;
    philox_u32_kernel<<<grid_for((n + 3) / 4), kThreads, 0, (cudaStream_t)stream>>>(seed, offset, n, d_out);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_philox_normal_ex(uint64_t seed, uint64_t offset, int64_t n, void *d_out, int out_dtype, int kernel_dtype, void *stream) {
    MLVAE_REQUIRE(d_out && n >= 0, MLVAE_ERR_INVALID_ARG, "philox_normal: bad arguments");
    MLVAE_REQUIRE((out_dtype == MLVAE_F32 || out_dtype == MLVAE_BF16) && (kernel_dtype == MLVAE_F32 || kernel_dtype == MLVAE_BF16), MLVAE_ERR_INVALID_ARG,
                  "philox_normal: unknown dtype");
    if (n == 0) return MLVAE_OK;
    const int g = grid_for((n + 3) / 4);
    cudaStream_t st = (cudaStream_t)stream;
    const bool v2 = kernel_dtype == MLVAE_BF16;
    if (out_dtype == MLVAE_F32) {
        if (v2) philox_normal_kernel<float, true><<<g, kThreads, 0, st>>>(seed, offset, n, (float *)d_out);
        else philox_normal_kernel<float, false><<<g, kThreads, 0, st>>>(seed, offset, n, (float *)d_out);
    } else {
        if (v2) philox_normal_kernel<__nv_bfloat16, true><<<g, kThreads, 0, st>>>(seed, offset, n, (__nv_bfloat16 *)d_out);
        else philox_normal_kernel<__nv_bfloat16, false><<<g, kThreads, 0, st>>>(seed, offset, n, (__nv_bfloat16 *)d_out);
    }
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_philox_normal(uint64_t seed, uint64_t offset, int64_t n, void *d_out, int dtype, void *stream) {
    return mlvae_philox_normal_ex(seed, offset, n, d_out, dtype, dtype, stream);      // the stream the kernels of that dtype draw
}

int mlvae_reparam_kl_fwd_strided(const void *d_mu, const void *d_logvar, int64_t ld_in, const void *d_eps, uint64_t seed, uint64_t offset,
                                 const uint64_t *d_offset_add, const float *d_lens, int B, int T_, int L, int dtype, void *d_z, void *d_kl_elem,
                                 float *d_kl_out, void *d_scratch, void *stream) {
    if (int rc = check_btc(B, T_, L)) return rc;
    MLVAE_REQUIRE(d_mu && d_logvar && d_z, MLVAE_ERR_INVALID_ARG, "reparam_kl_fwd: mu, logvar and z are required");
    MLVAE_REQUIRE(!d_kl_out || (d_lens && d_scratch), MLVAE_ERR_INVALID_ARG, "reparam_kl_fwd: reduced KL needs lens and scratch");
    if (ld_in == 0) ld_in = L;
    MLVAE_REQUIRE(ld_in >= L, MLVAE_ERR_INVALID_ARG, "reparam_kl_fwd: row stride %lld < L = %d", (long long)ld_in, L);
    const int esz = dtype == MLVAE_F32 ? 4 : 2;
    const bool al = aligned16(d_mu) && aligned16(d_logvar) && aligned16(d_z) && aligned16(d_eps) && aligned16(d_kl_elem) && (ld_in * esz) % 16 == 0;
    MLVAE_DISPATCH(dtype, L, al, {
        const int64_t nvec = (int64_t)B * T_ * (L / VEC);
        reparam_kl_fwd_kernel<T, VEC><<<grid_for(nvec), kThreads, 0, (cudaStream_t)stream>>>(
            (const T *)d_mu, (const T *)d_logvar, (const T *)d_eps, PhiloxKey(seed), offset, d_offset_add, d_lens, B, T_, L, ld_in, (T *)d_z,
            (T *)d_kl_elem, d_kl_out, (ReduceScratch *)d_scratch);
    });
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_reparam_kl_fwd(const void *d_mu, const void *d_logvar, const void *d_eps, uint64_t seed, uint64_t offset,
                         const uint64_t *d_offset_add, const float *d_lens, int B, int T_, int L, int dtype, void *d_z, void *d_kl_elem,
                         float *d_kl_out, void *d_scratch, void *stream) {
    return mlvae_reparam_kl_fwd_strided(d_mu, d_logvar, 0, d_eps, seed, offset, d_offset_add, d_lens, B, T_, L, dtype, d_z, d_kl_elem, d_kl_out,
                                        d_scratch, stream);
}

int mlvae_reparam_kl_bwd_strided(const void *d_mu, const void *d_logvar, int64_t ld_in, const void *d_eps, uint64_t seed, uint64_t offset,
                                 const uint64_t *d_offset_add, const void *d_grad_z, const void *d_grad_kl_elem, const float *d_grad_kl_mean,
                                 const float *d_lens, int B, int T_, int L, int dtype, void *d_grad_mu, void *d_grad_logvar, int64_t ld_out,
                                 void *stream) {
    if (int rc = check_btc(B, T_, L)) return rc;
    MLVAE_REQUIRE(d_mu && d_logvar && d_grad_mu && d_grad_logvar, MLVAE_ERR_INVALID_ARG, "reparam_kl_bwd: missing buffers");
    MLVAE_REQUIRE(!d_grad_kl_mean || d_lens, MLVAE_ERR_INVALID_ARG, "reparam_kl_bwd: reduced-KL gradient needs lens");
    if (ld_in == 0) ld_in = L;
    if (ld_out == 0) ld_out = L;
    MLVAE_REQUIRE(ld_in >= L && ld_out >= L, MLVAE_ERR_INVALID_ARG, "reparam_kl_bwd: row strides must be >= L = %d", L);
    const int esz = dtype == MLVAE_F32 ? 4 : 2;
    const bool al = aligned16(d_mu) && aligned16(d_logvar) && aligned16(d_eps) && aligned16(d_grad_z) &&
                    aligned16(d_grad_kl_elem) && aligned16(d_grad_mu) && aligned16(d_grad_logvar) && (ld_in * esz) % 16 == 0 && (ld_out * esz) % 16 == 0;
    MLVAE_DISPATCH(dtype, L, al, {
        const int64_t nvec = (int64_t)B * T_ * (L / VEC);
        reparam_kl_bwd_kernel<T, VEC><<<grid_for(nvec), kThreads, 0, (cudaStream_t)stream>>>(
            (const T *)d_mu, (const T *)d_logvar, (const T *)d_eps, PhiloxKey(seed), offset, d_offset_add, (const T *)d_grad_z,
            (const T *)d_grad_kl_elem, d_grad_kl_mean, d_lens, B, T_, L, ld_in, ld_out, (T *)d_grad_mu, (T *)d_grad_logvar);
    });
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_reparam_kl_bwd(const void *d_mu, const void *d_logvar, const void *d_eps, uint64_t seed, uint64_t offset,
                         const uint64_t *d_offset_add, const void *d_grad_z, const void *d_grad_kl_elem, const float *d_grad_kl_mean,
                         const float *d_lens, int B, int T_, int L, int dtype, void *d_grad_mu, void *d_grad_logvar,
                         void *stream) {
    return mlvae_reparam_kl_bwd_strided(d_mu, d_logvar, 0, d_eps, seed, offset, d_offset_add, d_grad_z, d_grad_kl_elem, d_grad_kl_mean, d_lens, B, T_, L,
                                        dtype, d_grad_mu, d_grad_logvar, 0, stream);
}

int mlvae_recon_fwd(const void *d_mean, const void *d_logvar, const void *d_target, const float *d_lens, int B, int T_,
                    int D, int dtype, int loss_type, void *d_elem, float *d_out, void *d_scratch, void *stream) {
    if (int rc = check_btc(B, T_, D)) return rc;
    MLVAE_REQUIRE(loss_type == MLVAE_RECON_LIKELIHOOD || loss_type == MLVAE_RECON_MSE, MLVAE_ERR_INVALID_ARG,
                  "Invalid loss type: %d", loss_type);
    const bool mse = loss_type == MLVAE_RECON_MSE;
    MLVAE_REQUIRE(d_mean && d_target && (mse || d_logvar), MLVAE_ERR_INVALID_ARG, "recon_fwd: missing inputs");
    MLVAE_REQUIRE(d_elem || d_out, MLVAE_ERR_INVALID_ARG, "recon_fwd: nothing to write");
    MLVAE_REQUIRE(!d_out || (d_lens && d_scratch), MLVAE_ERR_INVALID_ARG, "recon_fwd: reduced loss needs lens and scratch");
    const bool al = aligned16(d_mean) && aligned16(d_logvar) && aligned16(d_target) && aligned16(d_elem);
    MLVAE_DISPATCH(dtype, D, al, {
        const int64_t nvec = (int64_t)B * T_ * (D / VEC);
        const int g = grid_for(nvec);
        if (mse)
            recon_fwd_kernel<T, VEC, true><<<g, kThreads, 0, (cudaStream_t)stream>>>(
                (const T *)d_mean, (const T *)d_logvar, (const T *)d_target, d_lens, B, T_, D, (T *)d_elem, d_out,
                (ReduceScratch *)d_scratch);
        else
            recon_fwd_kernel<T, VEC, false><<<g, kThreads, 0, (cudaStream_t)stream>>>(
                (const T *)d_mean, (const T *)d_logvar, (const T *)d_target, d_lens, B, T_, D, (T *)d_elem, d_out,
                (ReduceScratch *)d_scratch);
    });
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_recon_bwd(const void *d_mean, const void *d_logvar, const void *d_target, const void *d_grad_elem,
                    const float *d_grad_mean_scalar, const float *d_lens, int B, int T_, int D, int dtype,
                    int loss_type, void *d_grad_mean, void *d_grad_logvar, void *d_grad_target, void *stream) {
    if (int rc = check_btc(B, T_, D)) return rc;
    MLVAE_REQUIRE(loss_type == MLVAE_RECON_LIKELIHOOD || loss_type == MLVAE_RECON_MSE, MLVAE_ERR_INVALID_ARG,
                  "Invalid loss type: %d", loss_type);
    const bool mse = loss_type == MLVAE_RECON_MSE;
    MLVAE_REQUIRE(d_mean && d_target && d_grad_mean && (mse || (d_logvar && d_grad_logvar)), MLVAE_ERR_INVALID_ARG,
                  "recon_bwd: missing buffers");
    MLVAE_REQUIRE(!d_grad_mean_scalar || d_lens, MLVAE_ERR_INVALID_ARG, "recon_bwd: reduced-loss gradient needs lens");
    const bool al = aligned16(d_mean) && aligned16(d_logvar) && aligned16(d_target) && aligned16(d_grad_elem) &&
                    aligned16(d_grad_mean) && aligned16(d_grad_logvar) && aligned16(d_grad_target);
    MLVAE_DISPATCH(dtype, D, al, {
        const int64_t nvec = (int64_t)B * T_ * (D / VEC);
        const int g = grid_for(nvec);
        if (mse)
            recon_bwd_kernel<T, VEC, true><<<g, kThreads, 0, (cudaStream_t)stream>>>(
                (const T *)d_mean, (const T *)d_logvar, (const T *)d_target, (const T *)d_grad_elem, d_grad_mean_scalar,
                d_lens, B, T_, D, (T *)d_grad_mean, (T *)d_grad_logvar, (T *)d_grad_target);
        else
            recon_bwd_kernel<T, VEC, false><<<g, kThreads, 0, (cudaStream_t)stream>>>(
                (const T *)d_mean, (const T *)d_logvar, (const T *)d_target, (const T *)d_grad_elem, d_grad_mean_scalar,
                d_lens, B, T_, D, (T *)d_grad_mean, (T *)d_grad_logvar, (T *)d_grad_target);
    });
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_gmm_reparam_kl_fwd(const void *d_mu, const void *d_logvar, const void *d_prior_mu, const void *d_prior_logvar,
                             const void *d_eps, uint64_t seed, uint64_t offset, const uint64_t *d_offset_add, int64_t n,
                             int dtype, void *d_z, void *d_kl_elem, void *stream) {
    MLVAE_REQUIRE(d_mu && d_logvar && d_prior_mu && d_prior_logvar && d_z && d_kl_elem && n > 0, MLVAE_ERR_INVALID_ARG,
                  "gmm_reparam_kl_fwd: missing buffers");
    const bool al = aligned16(d_mu) && aligned16(d_logvar) && aligned16(d_prior_mu) && aligned16(d_prior_logvar) &&
                    aligned16(d_eps) && aligned16(d_z) && aligned16(d_kl_elem);
    MLVAE_DISPATCH(dtype, n, al, {
        gmm_reparam_kl_fwd_kernel<T, VEC><<<grid_for(n / VEC), kThreads, 0, (cudaStream_t)stream>>>(
            (const T *)d_mu, (const T *)d_logvar, (const T *)d_prior_mu, (const T *)d_prior_logvar, (const T *)d_eps, seed, offset,
            d_offset_add, n, (T *)d_z, (T *)d_kl_elem);
    });
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_gmm_reparam_kl_bwd(const void *d_mu, const void *d_logvar, const void *d_prior_mu, const void *d_prior_logvar,
                             const void *d_eps, uint64_t seed, uint64_t offset, const uint64_t *d_offset_add,
                             const void *d_grad_z, const void *d_grad_kl_elem, int64_t n, int dtype, void *d_grad_mu,
                             void *d_grad_logvar, void *d_grad_prior_mu, void *d_grad_prior_logvar, void *stream) {
    MLVAE_REQUIRE(d_mu && d_logvar && d_prior_mu && d_prior_logvar && d_grad_mu && d_grad_logvar && d_grad_prior_mu &&
                      d_grad_prior_logvar && n > 0, MLVAE_ERR_INVALID_ARG, "gmm_reparam_kl_bwd: missing buffers");
    const bool al = aligned16(d_mu) && aligned16(d_logvar) && aligned16(d_prior_mu) && aligned16(d_prior_logvar) &&
                    aligned16(d_eps) && aligned16(d_grad_z) && aligned16(d_grad_kl_elem) && aligned16(d_grad_mu) &&
                    aligned16(d_grad_logvar) && aligned16(d_grad_prior_mu) && aligned16(d_grad_prior_logvar);
    MLVAE_DISPATCH(dtype, n, al, {
        gmm_reparam_kl_bwd_kernel<T, VEC><<<grid_for(n / VEC), kThreads, 0, (cudaStream_t)stream>>>(
            (const T *)d_mu, (const T *)d_logvar, (const T *)d_prior_mu, (const T *)d_prior_logvar, (const T *)d_eps, seed, offset,
            d_offset_add, (const T *)d_grad_z, (const T *)d_grad_kl_elem, n, (T *)d_grad_mu, (T *)d_grad_logvar,
            (T *)d_grad_prior_mu, (T *)d_grad_prior_logvar);
    });
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_apply_weight_fwd(const void *d_x, const void *d_w, int64_t M, int N, int C, int dtype, void *d_out, void *stream) {
    MLVAE_REQUIRE(d_x && d_w && d_out && M > 0 && N > 0 && C > 0, MLVAE_ERR_INVALID_ARG, "apply_weight_fwd: bad arguments");
    if (aligned16(d_x) && aligned16(d_out) && (dtype == MLVAE_F32 ? C % 4 == 0 : dtype == MLVAE_BF16 && C % 8 == 0)) {
        MLVAE_DISPATCH(dtype, C, true, {
            apply_weight_fwd_vec_kernel<T, VEC><<<grid_for(M * (C / VEC)), kThreads, 0, (cudaStream_t)stream>>>(
                (const T *)d_x, (const T *)d_w, M, N, C, (T *)d_out);
        });
        MLVAE_CHECK_CUDA(cudaGetLastError());
        return MLVAE_OK;
    }
    const int grid = grid_for(M * 32);
    if (dtype == MLVAE_F32) apply_weight_fwd_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const float *)d_x, (const float *)d_w, M, N, C, (float *)d_out);
    else if (dtype == MLVAE_BF16) apply_weight_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)d_x, (const __nv_bfloat16 *)d_w, M, N, C, (__nv_bfloat16 *)d_out);
    else return fail(MLVAE_ERR_INVALID_ARG, "unknown dtype %d", dtype);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_apply_weight_bwd(const void *d_x, const void *d_w, const void *d_grad_out, int64_t M, int N, int C, int dtype,
                           void *d_grad_x, void *d_grad_w, void *stream) {
    MLVAE_REQUIRE(d_x && d_w && d_grad_out && M > 0 && N > 0 && C > 0, MLVAE_ERR_INVALID_ARG, "apply_weight_bwd: bad arguments");
    {
        const int vec = dtype == MLVAE_F32 ? 4 : 8;
        const int vcn = C / vec;
        if ((dtype == MLVAE_F32 || dtype == MLVAE_BF16) && C % vec == 0 && vcn <= 32 && (vcn & (vcn - 1)) == 0 && aligned16(d_x) &&
            aligned16(d_grad_out) && (!d_grad_x || aligned16(d_grad_x))) {
            MLVAE_DISPATCH(dtype, C, true, {
                apply_weight_bwd_vec_kernel<T, VEC><<<grid_for(M * (C / VEC)), kThreads, 0, (cudaStream_t)stream>>>(
                    (const T *)d_x, (const T *)d_w, (const T *)d_grad_out, M, N, C, (T *)d_grad_x, (T *)d_grad_w);
            });
            MLVAE_CHECK_CUDA(cudaGetLastError());
            return MLVAE_OK;
        }
    }
    const int grid = grid_for(M * 32);
    if (dtype == MLVAE_F32) apply_weight_bwd_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const float *)d_x, (const float *)d_w, (const float *)d_grad_out, M, N, C, (float *)d_grad_x, (float *)d_grad_w);
    else if (dtype == MLVAE_BF16) apply_weight_bwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)d_x, (const __nv_bfloat16 *)d_w, (const __nv_bfloat16 *)d_grad_out, M, N, C, (__nv_bfloat16 *)d_grad_x, (__nv_bfloat16 *)d_grad_w);
    else return fail(MLVAE_ERR_INVALID_ARG, "unknown dtype %d", dtype);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_masked_reduce_fwd(const void *d_loss, const float *d_lens, int B, int T_, int C, int dtype, int reduction,
                            float *d_out, void *d_scratch, void *stream) {
    if (int rc = check_btc(B, T_, C)) return rc;
    MLVAE_REQUIRE(d_loss && d_lens && d_out, MLVAE_ERR_INVALID_ARG, "masked_reduce_fwd: missing buffers");
    MLVAE_REQUIRE(reduction >= MLVAE_RED_MEAN && reduction <= MLVAE_RED_BATCH, MLVAE_ERR_INVALID_ARG, "bad reduction %d", reduction);
    if (reduction == MLVAE_RED_BATCH) {
        if (dtype == MLVAE_F32) masked_rowmean_kernel<float><<<B, kThreads, 0, (cudaStream_t)stream>>>((const float *)d_loss, d_lens, T_, C, d_out);
        else if (dtype == MLVAE_BF16) masked_rowmean_kernel<__nv_bfloat16><<<B, kThreads, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)d_loss, d_lens, T_, C, d_out);
        else return fail(MLVAE_ERR_INVALID_ARG, "unknown dtype %d", dtype);
    } else {
        MLVAE_REQUIRE(d_scratch, MLVAE_ERR_INVALID_ARG, "masked_reduce_fwd: scratch required");
        MLVAE_DISPATCH(dtype, C, aligned16(d_loss), {
            const int64_t nvec = (int64_t)B * T_ * (C / VEC);
            masked_sum_kernel<T, VEC><<<grid_for(nvec), kThreads, 0, (cudaStream_t)stream>>>(
                (const T *)d_loss, d_lens, B, T_, C, reduction, d_out, (ReduceScratch *)d_scratch);
        });
    }
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_masked_reduce_bwd(const float *d_grad_out, const float *d_lens, int B, int T_, int C, int dtype, int reduction,
                            void *d_grad_loss, void *stream) {
    if (int rc = check_btc(B, T_, C)) return rc;
    MLVAE_REQUIRE(d_grad_out && d_lens && d_grad_loss, MLVAE_ERR_INVALID_ARG, "masked_reduce_bwd: missing buffers");
    MLVAE_REQUIRE(reduction >= MLVAE_RED_MEAN && reduction <= MLVAE_RED_BATCH, MLVAE_ERR_INVALID_ARG, "bad reduction %d", reduction);
    MLVAE_DISPATCH(dtype, C, aligned16(d_grad_loss), {
        const int64_t nvec = (int64_t)B * T_ * (C / VEC);
        masked_reduce_bwd_kernel<T, VEC><<<grid_for(nvec), kThreads, 0, (cudaStream_t)stream>>>(
            d_grad_out, d_lens, B, T_, C, reduction, (T *)d_grad_loss);
    });
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

}  // extern "C"
