// Backward prologue of Linear(+LeakyReLU) (reference: modules/fc_block.py:9-16 under autograd):
//   g  = dy * (y > 0 ? 1 : slope)        (only when the layer had the activation)
//   db = sum over rows of g              (bias gradient, float32, deterministic)
// in ONE pass over dy (and y).  torch needs where + cast + mul + a column-sum reduce_kernel that reaches
// ~0.4 TB/s on these (32000 x 64..128) shapes; this kernel reads every byte once with 16-byte vectors.
// Reduction: per-thread registers over rows -> shared memory over the CTA's row groups -> per-CTA partial
// -> fixed-order sum by the last-arriving CTA (self-resetting ticket, no float atomics).
#include "common.cuh"

namespace mlvae {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kDbThreads = 256;
__device__ __forceinline__ uint4 ld_stream(const __nv_bfloat16 *p) { return __ldcs(reinterpret_cast<const uint4 *>(p)); }
constexpr int kDbMaxGrid = 148 * 2;
constexpr int kDbUnroll = 4;             // rows in flight per thread (bytes in flight ~ latency x bandwidth)

struct DbScratch {
    unsigned int ticket;
    unsigned int pad[3];
    float partial[1];                 // [grid][N]
};

__global__ void __launch_bounds__(kDbThreads) dense_bwd_prep_kernel(const bf16 *__restrict__ dy, const bf16 *__restrict__ y,
                                                                    bf16 *__restrict__ g, float *__restrict__ db, int64_t M, int N,
                                                                    int64_t ld, float slope, DbScratch *scratch, int accumulate) {
    extern __shared__ float s_acc[];                     // [RP][N]
    __shared__ bool s_last;
    const int VC = N >> 3;                               // 16-byte vectors per row
    const int RP = kDbThreads / VC;                      // rows per pass
    const int tid = threadIdx.x;
    const int rl = tid / VC, vc = tid - rl * VC;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto consume = [&](uint4 d, uint4 a, size_t off) {
        uint32_t dw[4] = {d.x, d.y, d.z, d.w};
        if (y) {
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float lo = __uint_as_float(dw[e] << 16), hi = __uint_as_float(dw[e] & 0xffff0000u);
                if (!(__uint_as_float(aw[e] << 16) > 0.f)) lo *= slope;
                if (!(__uint_as_float(aw[e] & 0xffff0000u) > 0.f)) hi *= slope;
                const __nv_bfloat162 r = __floats2bfloat162_rn(lo, hi);
                dw[e] = *reinterpret_cast<const uint32_t *>(&r);
            }
            *reinterpret_cast<uint4 *>(g + off) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {                     // the bias gradient sums the ROUNDED g, like the weight GEMM that follows
            acc[2 * e] += __uint_as_float(dw[e] << 16);
            acc[2 * e + 1] += __uint_as_float(dw[e] & 0xffff0000u);
        }
    };
    if (rl < RP) {
        const int64_t stride = (int64_t)gridDim.x * RP;
        int64_t row = (int64_t)blockIdx.x * RP + rl;
        for (; row + (kDbUnroll - 1) * stride < M; row += kDbUnroll * stride) {
            uint4 d[kDbUnroll], a[kDbUnroll];
#pragma unroll
            for (int k = 0; k < kDbUnroll; ++k) {
                const size_t off = (size_t)(row + k * stride) * ld + (size_t)vc * 8;
                d[k] = ld_stream(dy + off);
                a[k] = y ? ld_stream(y + off) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < kDbUnroll; ++k) consume(d[k], a[k], (size_t)(row + k * stride) * ld + (size_t)vc * 8);
        }
        for (; row < M; row += stride) {
            const size_t off = (size_t)row * ld + (size_t)vc * 8;
            consume(ld_stream(dy + off), y ? ld_stream(y + off) : make_uint4(0, 0, 0, 0), off);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) s_acc[(size_t)rl * N + vc * 8 + e] = acc[e];
    }
    __syncthreads();
    float *mine = scratch->partial + (size_t)blockIdx.x * N;
    for (int c = tid; c < N; c += kDbThreads) {
        float s = 0.f;
        for (int r = 0; r < RP; ++r) s += s_acc[(size_t)r * N + c];
        mine[c] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&scratch->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // fixed-order sum of the per-CTA partials: P threads per column, each a contiguous run of CTAs, combined in order
    const int P = N <= kDbThreads ? kDbThreads / N : 1;
    const int per = ((int)gridDim.x + P - 1) / P;
    __syncthreads();
    for (int c0 = 0; c0 < N; c0 += kDbThreads) {
        const int c = c0 + (N <= kDbThreads ? tid % N : tid), part = N <= kDbThreads ? tid / N : 0;
        float s = 0.f;
        if (c < N && part < P) {
            const int b_end = min((int)gridDim.x, (part + 1) * per);
            int b = part * per;
            for (; b + 8 <= b_end; b += 8) {
                float v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = __ldcg(scratch->partial + (size_t)(b + k) * N + c);
#pragma unroll
                for (int k = 0; k < 8; ++k) s += v[k];
            }
            for (; b < b_end; ++b) s += __ldcg(scratch->partial + (size_t)b * N + c);
            s_acc[part * N + (c - c0)] = s;
        }
        __syncthreads();
        if (tid < N - c0 && tid < kDbThreads) {
            float t = 0.f;
            for (int q = 0; q < P; ++q) t += s_acc[q * N + tid];
            db[c0 + tid] = accumulate ? db[c0 + tid] + t : t;
        }
        __syncthreads();
    }
    if (tid == 0) scratch->ticket = 0;
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

extern "C" {

size_t mlvae_dense_bwd_scratch_bytes(int N) { return N > 0 ? sizeof(DbScratch) + (size_t)kDbMaxGrid * N * sizeof(float) : 0; }

int mlvae_dense_bwd_prep(const void *d_dy, const void *d_y, void *d_g, float *d_db, int64_t M, int N, int64_t ld, float slope,
                         void *d_scratch, int accumulate, void *stream) {
    MLVAE_REQUIRE(d_dy && d_db && d_scratch, MLVAE_ERR_INVALID_ARG, "dense_bwd_prep: missing buffers");
    MLVAE_REQUIRE((d_y == nullptr) == (d_g == nullptr), MLVAE_ERR_INVALID_ARG, "dense_bwd_prep: y and g go together");
    MLVAE_REQUIRE(M > 0 && N > 0 && N % 8 == 0 && N <= 2048 && ld >= N && ld % 8 == 0, MLVAE_ERR_UNSUPPORTED,
                  "dense_bwd_prep: needs N %% 8 == 0, N <= 2048 and 16-byte aligned rows (N=%d, ld=%lld)", N, (long long)ld);
    const int VC = N / 8, RP = kDbThreads / VC;
    MLVAE_REQUIRE(RP >= 1, MLVAE_ERR_UNSUPPORTED, "dense_bwd_prep: N too large");
    // enough CTAs to keep ~kDbUnroll rows in flight per thread, no more: every extra CTA is another partial for the last one to sum
    int64_t blocks = (M + (int64_t)RP * kDbUnroll - 1) / ((int64_t)RP * kDbUnroll);
    const int grid = (int)(blocks < 1 ? 1 : blocks < kDbMaxGrid ? blocks : kDbMaxGrid);
    const size_t smem = (size_t)RP * N * sizeof(float);
    dense_bwd_prep_kernel<<<grid, kDbThreads, smem, (cudaStream_t)stream>>>((const bf16 *)d_dy, (const bf16 *)d_y, (bf16 *)d_g, d_db, M,
                                                                             N, ld, slope, (DbScratch *)d_scratch, accumulate);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

}  // extern "C"
