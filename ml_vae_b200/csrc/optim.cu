// Optimiser step of the training hot path over the FLAT parameter arena (train_step.FlatArena), two launches:
//   check_gradients [SpeechBrain Brain, called at models/md_model.py:82]: clip the global gradient norm to max_grad_norm
//                   and skip the update when the loss is not finite
//   optimizer.step  [models/md_model.py:84-85 with torch.optim.Adam, models/test_vanilla_vae/model.yaml:45-47]
//   zero_grad       [models/md_model.py:86-87]
// plus what the data-parallel step needs around them: the 1 / world_size scale of the all-reduced gradient sum and the bf16
// shadow copy of the updated parameters that the tensor-core kernels read (no per-layer cast kernels in the next step).
// torch needs a foreach-norm, a foreach-mul, the fused Adam and a fill for the same work (4 launches, 2 extra passes over the
// gradients); here the gradients are read twice and everything else once: 8 x 4 bytes per parameter, HBM bound.
//
// Arithmetic follows torch.optim.Adam (fused, capturable) exactly, in float32:
//   step += 1;  m += (g - m) (1 - b1);  v = b2 v + (1 - b2) g g;
//   p -= (lr / (1 - b1^step)) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)
// with the hyper-parameters taken as DOUBLES like torch's (1 - b2 is rounded to float32 once from the double difference:
// 1.f - 0.999f would be off by 1.3e-5 relative)
// and torch.nn.utils.clip_grad_norm_:  g *= min(1, max_norm / (||g||_2 + 1e-6)).
// Deterministic: per-CTA partial sums of squares, added in index order by every CTA of the second kernel.
#include "common.cuh"

namespace mlvae {
namespace {

constexpr int kOptThreads = 256;
constexpr int kOptMaxGrid = 148 * 8;

struct AdamState {
    float step;          // number of updates applied so far (float like torch's capturable step tensor)
    float last_norm;     // gradient norm of the last call (after the all-reduce scale, before clipping)
    float last_coef;     // clip coefficient of the last call
    float pad;
    float partial[kOptMaxGrid];
};

__global__ void __launch_bounds__(kOptThreads) grad_sumsq_kernel(const float *__restrict__ g, int64_t n, float gscale, AdamState *st,
                                                                const float *__restrict__ loss) {
    float acc = 0.f;
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kOptThreads) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(g) + i);
        const float a = v.x * gscale, b = v.y * gscale, c = v.z * gscale, d = v.w * gscale;
        acc += (a * a + b * b) + (c * c + d * d);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t i = n4 << 2; i < n; ++i) acc += (g[i] * gscale) * (g[i] * gscale);
    const float s = block_sum(acc);
    if (threadIdx.x == 0) {
        st->partial[blockIdx.x] = s;
        if (blockIdx.x == 0 && !(loss && !isfinite(*loss))) st->step += 1.f;      // the update will be applied
    }
}

__global__ void __launch_bounds__(kOptThreads) adam_step_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m,
                                                               float *__restrict__ v, __nv_bfloat16 *__restrict__ p16, int64_t n, int npart,
                                                               float gscale, double lr, double b1d, double b2d, float eps, float max_norm,
                                                               AdamState *st, const float *__restrict__ loss) {
    const float b2 = (float)b2d, omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d);
    __shared__ float s_coef;
    {   // every CTA adds the partials in the same order
        float a = 0.f;
        for (int i = threadIdx.x; i < npart; i += kOptThreads) a += st->partial[i];
        const float tot = block_sum(a);
        if (threadIdx.x == 0) {
            const float norm = sqrtf(tot);
            float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
            coef = fminf(coef, 1.f);
            s_coef = coef * gscale;
            if (blockIdx.x == 0) { st->last_norm = norm; st->last_coef = coef; }
        }
        __syncthreads();
    }
    const bool skip = loss && !isfinite(*loss);
    const float coef = s_coef;
    const float step = st->step;
    // bias corrections in double like torch's python-side arithmetic (once per thread)
    const double bc1 = 1.0 - pow(b1d, (double)step), bc2 = 1.0 - pow(b2d, (double)step);
    const float step_size = (float)(lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kOptThreads) {
        float4 pw = reinterpret_cast<float4 *>(p)[i];
        if (!skip) {
            const float4 gw = reinterpret_cast<const float4 *>(g)[i];
            float4 mw = reinterpret_cast<float4 *>(m)[i], vw = reinterpret_cast<float4 *>(v)[i];
            const float gg[4] = {gw.x * coef, gw.y * coef, gw.z * coef, gw.w * coef};
            float pp[4] = {pw.x, pw.y, pw.z, pw.w}, mm[4] = {mw.x, mw.y, mw.z, mw.w}, vv[4] = {vw.x, vw.y, vw.z, vw.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                mm[e] = mm[e] + (gg[e] - mm[e]) * omb1;
                vv[e] = vv[e] * b2 + omb2 * gg[e] * gg[e];
                const float denom = sqrtf(vv[e]) / bc2_sqrt + eps;
                pp[e] = pp[e] - step_size * (mm[e] / denom);
            }
            pw = make_float4(pp[0], pp[1], pp[2], pp[3]);
            reinterpret_cast<float4 *>(p)[i] = pw;
            reinterpret_cast<float4 *>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
            reinterpret_cast<float4 *>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        }
        reinterpret_cast<float4 *>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);                 // zero_grad
        if (p16) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(pw.x, pw.y), hi = __floats2bfloat162_rn(pw.z, pw.w);
            reinterpret_cast<uint2 *>(p16)[i] = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = n4 << 2; i < n; ++i) {
            if (!skip) {
                const float ge = g[i] * coef;
                m[i] = m[i] + (ge - m[i]) * omb1;
                v[i] = v[i] * b2 + omb2 * ge * ge;
                p[i] = p[i] - step_size * (m[i] / (sqrtf(v[i]) / bc2_sqrt + eps));
            }
            g[i] = 0.f;
            if (p16) p16[i] = __float2bfloat16_rn(p[i]);
        }
    }
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

extern "C" {

size_t mlvae_adam_state_bytes(void) { return sizeof(AdamState); }

int mlvae_adam_clip_step(float *d_params, float *d_grads, float *d_exp_avg, float *d_exp_avg_sq, void *d_params_bf16, int64_t n, float grad_scale,
                         double lr, double beta1, double beta2, double eps, float max_grad_norm, void *d_state, const float *d_loss, void *stream) {
    MLVAE_REQUIRE(d_params && d_grads && d_exp_avg && d_exp_avg_sq && d_state && n > 0, MLVAE_ERR_INVALID_ARG, "adam_clip_step: missing buffers");
    MLVAE_REQUIRE(((uintptr_t)d_params & 15) == 0 && ((uintptr_t)d_grads & 15) == 0 && ((uintptr_t)d_exp_avg & 15) == 0 &&
                      ((uintptr_t)d_exp_avg_sq & 15) == 0 && ((uintptr_t)d_params_bf16 & 7) == 0,
                  MLVAE_ERR_INVALID_ARG, "adam_clip_step: buffers must be 16-byte aligned");
    int64_t blocks = ((n >> 2) + kOptThreads - 1) / kOptThreads;
    const int64_t cap = (int64_t)sm_count() * 8 < kOptMaxGrid ? (int64_t)sm_count() * 8 : kOptMaxGrid;
    const int grid = (int)(blocks < 1 ? 1 : blocks > cap ? cap : blocks);
    cudaStream_t st = (cudaStream_t)stream;
    grad_sumsq_kernel<<<grid, kOptThreads, 0, st>>>(d_grads, n, grad_scale, (AdamState *)d_state, d_loss);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    adam_step_kernel<<<grid, kOptThreads, 0, st>>>(d_params, d_grads, d_exp_avg, d_exp_avg_sq, (__nv_bfloat16 *)d_params_bf16, n, grid, grad_scale, lr, beta1,
                                                  beta2, (float)eps, max_grad_norm, (AdamState *)d_state, d_loss);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

}  // extern "C"
