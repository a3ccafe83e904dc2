// Global input normalisation = speechbrain.processing.features.InputNormalization(norm_type='global') as declared at
// models/test_vanilla_vae/model.yaml:14-15 and called at models/test_vanilla_vae/model.py:24-25 [arithmetic: SB-recall,
// SpeechBrain 0.5.x]: per-utterance mean and unbiased std over the round(len*T) valid frames (std floored at 1e-10),
// averaged over the batch, folded into running statistics with weight 1/(count+1) while epoch < update_until_epoch,
// output (x - glob_mean) / glob_std.  SpeechBrain loops over the batch in python with an .int() sync per utterance;
// here the state (count, running mean / std) lives on the device, so the whole step stays capturable in a CUDA graph.
//   stats_kernel   one CTA per utterance: two-pass mean / variance per feature column over the valid frames
//   update_kernel  one CTA: batch average, running-average update
//   apply_kernel   streaming (x - mean) * (1/std) -> out dtype
#include "common.cuh"

namespace mlvae {
namespace {

constexpr int kNormThreads = 256;

// device-resident state: {count, pad[3], glob_mean[D], glob_std[D]} (float32)

__global__ void __launch_bounds__(1024)
norm_stats_kernel(const float *__restrict__ x, const float *__restrict__ lens, int T, int D, float eps,
                  float *__restrict__ mean_out, float *__restrict__ std_out) {
    // column c handled by threads {c, c + D, ...}: each strides over frames; partial sums combined through smem
    extern __shared__ float s_red[];            // [rows_per_pass][D]
    const int b = blockIdx.x;
    const float len = __ldg(lens + b);
    int n = (int)rintf(len * (float)T);         // torch.round(lengths * T).int(): round half to even, like rintf
    n = max(0, min(n, T));
    const float *xb = x + (size_t)b * T * D;
    const int groups = (int)blockDim.x / D > 0 ? (int)blockDim.x / D : 1;    // threads per column
    const int c = threadIdx.x % D, g = threadIdx.x / D;
    const bool active = threadIdx.x < groups * D;
    float s = 0.f;
    if (active)
        for (int t = g; t < n; t += groups) s += xb[(size_t)t * D + c];
    if (active) s_red[g * D + c] = s;
    __syncthreads();
    float mean = 0.f;
    if (active) {
        for (int k = 0; k < groups; ++k) mean += s_red[k * D + c];
        mean /= (float)n;
    }
    __syncthreads();
    float v = 0.f;
    if (active)
        for (int t = g; t < n; t += groups) {
            const float dlt = xb[(size_t)t * D + c] - mean;
            v = fmaf(dlt, dlt, v);
        }
    if (active) s_red[g * D + c] = v;
    __syncthreads();
    if (threadIdx.x < D) {
        float var = 0.f;
        for (int k = 0; k < groups; ++k) var += s_red[k * D + threadIdx.x];
        var /= (float)(n - 1);                   // unbiased, like torch.std (n == 1 -> NaN, as in the reference)
        mean_out[(size_t)b * D + threadIdx.x] = mean;
        std_out[(size_t)b * D + threadIdx.x] = fmaxf(sqrtf(var), eps);
    }
}

__global__ void __launch_bounds__(1024)
norm_update_kernel(const float *__restrict__ mean_b, const float *__restrict__ std_b, int B, int D, int update,
                   float *__restrict__ state, float scale, float *__restrict__ avg_out) {
    // scale: 1 / B (batch average) or 1 / world (rows = 1: the all-reduced sum of every rank's batch average).
    // avg_out != nullptr: only write the average {mean[D], std[D]} there (data-parallel statistics exchange), leave the state alone.
    // state = {count, pad[3], glob_mean[D], glob_std[D]}.  Column c is summed by `groups` threads over interleaved
    // utterances, then combined in fixed group order (deterministic).
    extern __shared__ float s_red[];            // [2][groups][D]
    float *gm = state + 4, *gs = state + 4 + D;
    const float count = state[0];
    const int groups = (int)blockDim.x / D > 0 ? (int)blockDim.x / D : 1;
    const int c = threadIdx.x % D, g = threadIdx.x / D;
    if (g < groups && threadIdx.x < groups * D) {
        float m = 0.f, sd = 0.f;
        for (int b = g; b < B; b += groups) { m += mean_b[(size_t)b * D + c]; sd += std_b[(size_t)b * D + c]; }
        s_red[g * D + c] = m;
        s_red[(groups + g) * D + c] = sd;
    }
    __syncthreads();
    for (int col = threadIdx.x; col < D; col += blockDim.x) {
        float m = 0.f, sd = 0.f;
        for (int k = 0; k < groups; ++k) { m += s_red[k * D + col]; sd += s_red[(groups + k) * D + col]; }
        m *= scale; sd *= scale;
        if (avg_out) { avg_out[col] = m; avg_out[D + col] = sd; continue; }
        if (count == 0.f) { gm[col] = m; gs[col] = sd; }
        else if (update) {
            const float w = 1.f / (count + 1.f);
            gm[col] = (1.f - w) * gm[col] + w * m;
            gs[col] = (1.f - w) * gs[col] + w * sd;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && !avg_out) state[0] = count + 1.f;
}

template <typename T>
__global__ void __launch_bounds__(kNormThreads)
norm_apply_kernel(const float *__restrict__ x, const float *__restrict__ state, int64_t rows, int D, T *__restrict__ out) {
    const float *gm = state + 4, *gs = state + 4 + D;
    const int64_t n = rows * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % D);
        out[i] = from_f32<T>((x[i] - gm[c]) / gs[c]);
    }
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

extern "C" {

// bytes of the device-resident normaliser state for D feature columns (zero it once: count = 0)
size_t mlvae_norm_state_bytes(int D) { return sizeof(float) * (4 + 2 * (size_t)D); }
size_t mlvae_norm_scratch_bytes(int B, int D) { return sizeof(float) * 2 * (size_t)B * D; }

// x (B,T,D) float32, lens (B,) relative lengths; training != 0 folds this batch into the running statistics
// (update_stats != 0 mirrors `epoch < update_until_epoch`; the very first batch always initialises them);
// out (B,T,D) of out_dtype = (x - glob_mean) / glob_std.
int mlvae_global_norm(const float *d_x, const float *d_lens, int B, int T, int D, int training, int update_stats,
                      float *d_state, float *d_scratch, void *d_out, int out_dtype, void *stream) {
    MLVAE_REQUIRE(d_x && d_lens && d_state && d_scratch && d_out, MLVAE_ERR_INVALID_ARG, "global_norm: missing buffers");
    MLVAE_REQUIRE(B > 0 && T > 0 && D > 0 && D <= 1024, MLVAE_ERR_INVALID_ARG, "global_norm: bad sizes (D <= 1024)");
    cudaStream_t st = (cudaStream_t)stream;
    if (training) {
        float *mean_b = d_scratch, *std_b = d_scratch + (size_t)B * D;
        const int threads = 1024;                                   // B CTAs only: as many frames in flight per CTA as possible
        const int groups = threads / D > 0 ? threads / D : 1;
        norm_stats_kernel<<<B, threads, sizeof(float) * groups * D, st>>>(d_x, d_lens, T, D, 1e-10f, mean_b, std_b);
        MLVAE_CHECK_CUDA(cudaGetLastError());
        norm_update_kernel<<<1, threads, sizeof(float) * 2 * groups * D, st>>>(mean_b, std_b, B, D, update_stats, d_state, 1.f / (float)B, nullptr);
        MLVAE_CHECK_CUDA(cudaGetLastError());
    }
    const int64_t n = (int64_t)B * T * D;
    int grid = (int)((n + kNormThreads - 1) / kNormThreads);
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    if (out_dtype == MLVAE_F32) norm_apply_kernel<float><<<grid, kNormThreads, 0, st>>>(d_x, d_state, (int64_t)B * T, D, (float *)d_out);
    else if (out_dtype == MLVAE_BF16) norm_apply_kernel<__nv_bfloat16><<<grid, kNormThreads, 0, st>>>(d_x, d_state, (int64_t)B * T, D, (__nv_bfloat16 *)d_out);
    else return fail(MLVAE_ERR_INVALID_ARG, "unknown dtype %d", out_dtype);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

// Data-parallel variant in two calls with the caller's all-reduce in between ("optional second tiny all-reduce for
// normalizer stats (2 D floats)"): every rank's batch contributes to ONE set of running statistics, as if the global batch had been seen.
//   1. mlvae_global_norm_batch_avg: per-utterance statistics of this rank's batch, averaged -> d_avg {mean[D], std[D]}
//   2. caller: all-reduce(sum) of d_avg over the ranks (equal batch sizes)
//   3. mlvae_global_norm_from_avg: running update with avg_scale * d_avg (avg_scale = 1 / world), then (x - glob_mean) / glob_std
int mlvae_global_norm_batch_avg(const float *d_x, const float *d_lens, int B, int T, int D, float *d_scratch, float *d_avg, void *stream) {
    MLVAE_REQUIRE(d_x && d_lens && d_scratch && d_avg, MLVAE_ERR_INVALID_ARG, "global_norm_batch_avg: missing buffers");
    MLVAE_REQUIRE(B > 0 && T > 0 && D > 0 && D <= 1024, MLVAE_ERR_INVALID_ARG, "global_norm_batch_avg: bad sizes (D <= 1024)");
    cudaStream_t st = (cudaStream_t)stream;
    float *mean_b = d_scratch, *std_b = d_scratch + (size_t)B * D;
    const int threads = 1024;
    const int groups = threads / D > 0 ? threads / D : 1;
    norm_stats_kernel<<<B, threads, sizeof(float) * groups * D, st>>>(d_x, d_lens, T, D, 1e-10f, mean_b, std_b);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    norm_update_kernel<<<1, threads, sizeof(float) * 2 * groups * D, st>>>(mean_b, std_b, B, D, 0, d_avg, 1.f / (float)B, d_avg);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

int mlvae_global_norm_from_avg(const float *d_x, int B, int T, int D, const float *d_avg, float avg_scale, int update_stats, float *d_state,
                               void *d_out, int out_dtype, void *stream) {
    MLVAE_REQUIRE(d_x && d_avg && d_state && d_out, MLVAE_ERR_INVALID_ARG, "global_norm_from_avg: missing buffers");
    MLVAE_REQUIRE(B > 0 && T > 0 && D > 0 && D <= 1024 && avg_scale > 0.f, MLVAE_ERR_INVALID_ARG, "global_norm_from_avg: bad sizes (D <= 1024) or scale");
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 1024;
    const int groups = threads / D > 0 ? threads / D : 1;
    norm_update_kernel<<<1, threads, sizeof(float) * 2 * groups * D, st>>>(d_avg, d_avg + D, 1, D, update_stats, d_state, avg_scale, nullptr);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    const int64_t n = (int64_t)B * T * D;
    int grid = (int)((n + kNormThreads - 1) / kNormThreads);
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    if (out_dtype == MLVAE_F32) norm_apply_kernel<float><<<grid, kNormThreads, 0, st>>>(d_x, d_state, (int64_t)B * T, D, (float *)d_out);
    else if (out_dtype == MLVAE_BF16) norm_apply_kernel<__nv_bfloat16><<<grid, kNormThreads, 0, st>>>(d_x, d_state, (int64_t)B * T, D, (__nv_bfloat16 *)d_out);
    else return fail(MLVAE_ERR_INVALID_ARG, "unknown dtype %d", out_dtype);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

}  // extern "C"
