// tcgen05 / TMEM dense projections for sm_100a.
//   mlvae_tc05_selftest : one 128 x N x K tile through the tensor-core path (descriptor / TMEM
//                         layout check used by tests/test_lstm_gpu.py::test_tcgen05_tile_matches_matmul)
//   mlvae_linear_fwd    : Y = act(X W^T + b), the building block of the FC stacks
//                         (modules/fc_block.py:4-21, vanilla_vae.py:22-24, decoder.py:24-25)
#include "common.cuh"
#include "tc05.cuh"

namespace mlvae {
namespace {

using bf16 = __nv_bfloat16;

// Copy a (rows x kdim) row-major bf16 matrix slice into the no-swizzle K-major tile layout.
// Consecutive lanes take consecutive rows (16-byte apart in the tile -> conflict-free stores).
__device__ __forceinline__ void load_tile_kmajor(unsigned char *tile, const bf16 *src, int64_t ld, int rows, int rows_valid,
                                                 int kdim, int tid, int nthreads) {
    const int chunks = kdim >> 3;                       // 16-byte chunks per row
    for (int i = tid; i < rows * chunks; i += nthreads) {
        const int r = (i & 7) | ((i / (8 * chunks)) << 3);
        const int c = (i >> 3) % chunks;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < rows_valid) v = __ldg(reinterpret_cast<const uint4 *>(src + (int64_t)r * ld + c * 8));
        *reinterpret_cast<uint4 *>(tile + tc::kmajor_off(r, c * 8, kdim)) = v;
    }
}

__global__ void __launch_bounds__(128, 1)
tc05_selftest_kernel(const bf16 *__restrict__ A, const bf16 *__restrict__ B, float *__restrict__ D, int N, int K, uint32_t tmem_cols,
                     int a_in_tmem) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    unsigned char *sA = smem;
    unsigned char *sB = smem + 128 * K * 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) tc::tmem_alloc(&s_tmem, tmem_cols);
    if (tid == 0) {
        tc::mbar_init(&s_bar, 1);
        tc::fence_barrier_init();
    }
    load_tile_kmajor(sA, A, K, 128, 128, K, tid, 128);
    load_tile_kmajor(sB, B, K, N, N, K, tid, 128);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t a_tmem = tmem + 256;          // A operand columns [256, 256 + K/2) when a_in_tmem

    if (a_in_tmem) {
        // thread = row m (lane 32*warp + lane): K bf16 of its row -> K/2 packed 32-bit TMEM columns
        const int m = warp * 32 + lane;
        for (int k0 = 0; k0 < K; k0 += 16) {
            const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(A + (int64_t)m * K + k0));
            const uint4 hi = __ldg(reinterpret_cast<const uint4 *>(A + (int64_t)m * K + k0 + 8));
            const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            tc::tmem_st8(a_tmem + ((uint32_t)(warp * 32) << 16) + k0 / 2, v);
        }
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_bf16_f32(128, N);
        const uint32_t sbo = (uint32_t)(K >> 3) * 128;
        for (int k = 0; k < K / 16; ++k) {
            const uint64_t da = tc::smem_desc(tc::smem_u32(sA) + k * 256, 128, sbo);
            const uint64_t db = tc::smem_desc(tc::smem_u32(sB) + k * 256, 128, sbo);
            if (a_in_tmem) tc::mma_bf16_ts(tmem, a_tmem + k * 8, db, idesc, k > 0);
            else tc::mma_bf16(tmem, da, db, idesc, k > 0);
        }
        tc::mma_commit(&s_bar);
    }
    tc::mbar_wait(&s_bar, 0);
    tc::fence_after_sync();
    const int row = warp * 32 + lane;
    for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) D[(int64_t)row * N + c + j] = __uint_as_float(v[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, tmem_cols);
}


// ======================================================================================
// Y[M, N] = act(X[M, K] W[N, K]^T + b[N])      bf16 in / bf16 out, fp32 accumulation in TMEM.
// One CTA = one 128-row tile of X.  K is consumed in chunks of 64 through a kStages-deep ring of
// shared-memory stages {X chunk 128 x 64, W chunk Npad x 64} filled with 16-byte cp.async (zero fill
// for rows >= M, W rows >= N and the K tail); one elected thread issues 4 tcgen05.mma per chunk and
// commits to the stage's mbarrier so the stage can be refilled; the epilogue reads the accumulator
// with tcgen05.ld (lane = row), adds the bias, applies LeakyReLU(0.01) if asked, converts and
// stores 16-byte vectors.  These projections are HBM-bound skinny GEMMs (N <= 256): the point of
// the kernel is one pass over X with no intermediate round trips, not tensor-pipe utilisation.
// ======================================================================================
constexpr int kLinThreads = 256;
constexpr int kBK = 64;
// Warp-specialised: warps 1..7 (224 threads) are cp.async producers running up to kStages chunks ahead (about 96 KB
// per SM must be in flight to cover the DRAM latency at full bandwidth); every producer thread arrives on the stage's
// "full" mbarrier through cp.async.mbarrier.arrive.noinc, i.e. when ITS copies have landed -- no wait_group, no block
// barrier, and no proxy fence in the producers (a fence.proxy.async there drains every in-flight cp.async and
// serialises the pipeline: measured 2400 cycles per chunk).  Warp 0 waits "full", fences generic->async proxy once,
// issues the 4 MMAs of the chunk and commits to the stage's "free" mbarrier.
constexpr int kProducers = kLinThreads - 32;

struct LinearParams {
    const bf16 *X;       // (M, K) row-major, ld = ldx
    const bf16 *W;       // (N, K) row-major
    const float *bias;   // (N) or nullptr
    bf16 *Y;             // (M, N) row-major, ld = ldy
    int M, N, K, ldx, ldy, leaky, stages;
};

constexpr int kMaxStages = 8;

__global__ void __launch_bounds__(kLinThreads, 1) linear_fwd_kernel(LinearParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // the 128-byte swizzle is a function of absolute shared-memory address bits [4,10): align the ring by hand
    unsigned char *smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t s_full[kMaxStages];  // chunk landed in the stage
    __shared__ uint64_t s_free[kMaxStages];  // stage may be overwritten (its MMAs completed)
    __shared__ uint64_t s_done;              // all MMAs of the tile completed
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Npad = (p.N + 15) & ~15;
    const int KC = (p.K + kBK - 1) / kBK;
    const int S = p.stages;
    const int m0 = blockIdx.x * 128;
    const size_t a_bytes = 128 * kBK * 2, w_bytes = (size_t)Npad * kBK * 2, st_bytes = a_bytes + w_bytes;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < Npad) tmem_cols <<= 1;

    if (warp == 0) tc::tmem_alloc(&s_tmem, tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < S; ++i) {
            tc::mbar_init(&s_full[i], kProducers);
            tc::mbar_init(&s_free[i], 1);
        }
        tc::mbar_init(&s_done, 1);
        tc::fence_barrier_init();
    }
    float *s_bias = reinterpret_cast<float *>(smem + (size_t)S * st_bytes);
    for (int i = tid; i < Npad; i += kLinThreads) s_bias[i] = (p.bias && i < p.N) ? __ldg(p.bias + i) : 0.f;
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;

    if (warp > 0) {
        // ================= producers =================
        const int ptid = tid - 32;
        for (int kc = 0; kc < KC; ++kc) {
            const int st = kc % S;
            if (kc >= S) tc::mbar_wait(&s_free[st], ((kc / S) - 1) & 1);
            unsigned char *sa = smem + (size_t)st * st_bytes, *sw = sa + a_bytes;
            const int k0 = kc * kBK;
            // 8 consecutive lanes copy one 128-byte row segment (fully coalesced lines); the 128-byte swizzle makes the
            // shared-memory side conflict free (each quarter-warp writes one whole row = all 32 banks)
            for (int i = ptid; i < 128 * 8; i += kProducers) {                  // X chunk: 128 rows x 8 16-byte pieces
                const int r = i >> 3, c = i & 7;
                const int m = m0 + r, k = k0 + c * 8;
                const bool ok = (m < p.M) && (k < p.K);
                tc::cp_async16(sa + tc::sw128_off(r, c), p.X + (size_t)(ok ? m : 0) * p.ldx + (ok ? k : 0), ok ? 16u : 0u);
            }
            for (int i = ptid; i < Npad * 8; i += kProducers) {                 // W chunk: Npad rows x 8 pieces
                const int r = i >> 3, c = i & 7;
                const int k = k0 + c * 8;
                const bool ok = (r < p.N) && (k < p.K);
                tc::cp_async16(sw + tc::sw128_off(r, c), p.W + (size_t)(ok ? r : 0) * p.K + (ok ? k : 0), ok ? 16u : 0u);
            }
            tc::cp_async_arrive_noinc(&s_full[st]);
        }
    } else {
        // ================= MMA issuer (warp 0, one elected lane) =================
        const uint32_t idesc = tc::idesc_bf16_f32(128, Npad);
        for (int kc = 0; kc < KC; ++kc) {
            const int st = kc % S;
            tc::mbar_wait(&s_full[st], (kc / S) & 1);
            tc::fence_proxy_async();                    // producers' generic-proxy writes -> async proxy (MMA operand reads)
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const uint32_t sa = tc::smem_u32(smem + (size_t)st * st_bytes), sw = sa + (uint32_t)a_bytes;
                const uint64_t da = tc::smem_desc_sw128(sa), dw = tc::smem_desc_sw128(sw);
#pragma unroll
                for (int k16 = 0; k16 < kBK / 16; ++k16)
                    tc::mma_bf16(tmem, da + (uint64_t)(k16 * 2), dw + (uint64_t)(k16 * 2), idesc, (kc | k16) != 0);
                tc::mma_commit(&s_free[st]);
                if (kc == KC - 1) tc::mma_commit(&s_done);
            }
            __syncwarp();
        }
    }
    tc::mbar_wait(&s_done, 0);
    tc::fence_after_sync();
    // ---- epilogue: warp (q = warp & 3) owns rows 32q..32q+31, column half (warp >> 2) ----
    const int q = warp & 3, half = warp >> 2;
    const int row = m0 + q * 32 + lane;
    const int ncol16 = Npad / 16;
    for (int cb = half; cb < ncol16; cb += 2) {
        uint32_t v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + cb * 16, v);
        tc::tmem_ld_wait();
        if (row < p.M) {
            uint32_t packed[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
                const int c0 = cb * 16 + 2 * j;
                a += s_bias[c0];
                b += s_bias[c0 + 1];
                if (p.leaky) {
                    a = a > 0.f ? a : 0.01f * a;
                    b = b > 0.f ? b : 0.01f * b;
                }
                const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                packed[j] = *reinterpret_cast<const uint32_t *>(&h);
            }
            bf16 *dst = p.Y + (size_t)row * p.ldy + cb * 16;
            if (cb * 16 + 16 <= p.N && (p.ldy % 8) == 0) {
                *reinterpret_cast<uint4 *>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                *reinterpret_cast<uint4 *>(dst + 8) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
            } else {
                const bf16 *src = reinterpret_cast<const bf16 *>(packed);
                for (int j = 0; j < 16; ++j)
                    if (cb * 16 + j < p.N) dst[j] = src[j];
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, tmem_cols);
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

extern "C" {

// D (128 x N, float32) = A (128 x K, bf16 row-major) * B (N x K, bf16 row-major)^T.  N % 16 == 0, N <= 256,
// K % 16 == 0, (128 + N) * K * 2 <= 200 KB.  Test hook for the descriptor conventions in tc05.cuh.
int mlvae_tc05_selftest(const void *d_a, const void *d_b, float *d_d, int N, int K, int a_in_tmem, void *stream) {
    MLVAE_REQUIRE(d_a && d_b && d_d, MLVAE_ERR_INVALID_ARG, "tc05_selftest: null buffer");
    MLVAE_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K % 16 == 0, MLVAE_ERR_INVALID_ARG, "tc05_selftest: bad N/K");
    const size_t smem = (size_t)(128 + N) * K * 2;
    MLVAE_REQUIRE(smem <= 200 * 1024, MLVAE_ERR_UNSUPPORTED, "tc05_selftest: tile too large");
    uint32_t cols = 32;
    while ((int)cols < N) cols <<= 1;
    if (a_in_tmem) {
        MLVAE_REQUIRE(K <= 512, MLVAE_ERR_UNSUPPORTED, "tc05_selftest: A-in-TMEM needs K <= 512");
        cols = 512;
    }
    MLVAE_CHECK_CUDA(cudaFuncSetAttribute(tc05_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc05_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const bf16 *)d_a, (const bf16 *)d_b, d_d, N, K, cols, a_in_tmem);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}


// Y (M x N, bf16, ld = ldy) = act(X (M x K, bf16, ld = ldx) W (N x K, bf16)^T + bias (N, f32 or NULL)); leaky != 0 applies
// LeakyReLU(0.01).  K % 8 == 0, ldx % 8 == 0, N <= 256, 16-byte aligned X / W.
int mlvae_linear_fwd(const void *d_x, const void *d_w, const float *d_bias, void *d_y, int M, int N, int K, int ldx, int ldy,
                     int leaky, void *stream) {
    MLVAE_REQUIRE(d_x && d_w && d_y, MLVAE_ERR_INVALID_ARG, "linear_fwd: null buffer");
    MLVAE_REQUIRE(M > 0 && N > 0 && K > 0, MLVAE_ERR_INVALID_ARG, "linear_fwd: bad sizes");
    MLVAE_REQUIRE(N <= 256 && K % 8 == 0 && ldx % 8 == 0 && ldx >= K && ldy >= N, MLVAE_ERR_UNSUPPORTED,
                  "linear_fwd: needs N <= 256, K %% 8 == 0, ldx %% 8 == 0 (got N=%d K=%d ldx=%d)", N, K, ldx);
    MLVAE_REQUIRE(((uintptr_t)d_x & 15) == 0 && ((uintptr_t)d_w & 15) == 0, MLVAE_ERR_INVALID_ARG, "linear_fwd: X and W must be 16-byte aligned");
    const int Npad = (N + 15) & ~15;
    const size_t st_bytes = 128 * kBK * 2 + (size_t)Npad * kBK * 2;
    const int KC = (K + kBK - 1) / kBK;
    int stages = (int)((128 * 1024 + st_bytes - 1) / st_bytes);        // ~128 KB of ring per SM
    if (stages > KC) stages = KC;
    if (stages < 2) stages = 2;
    if (stages > kMaxStages) stages = kMaxStages;
    while (stages > 2 && (size_t)stages * st_bytes + Npad * 4 + 1024 > 200 * 1024) --stages;
    const size_t smem = (size_t)stages * st_bytes + Npad * 4 + 1024;
    LinearParams prm{(const bf16 *)d_x, (const bf16 *)d_w, d_bias, (bf16 *)d_y, M, N, K, ldx, ldy, leaky, stages};
    // per device and cheap: set on every call (a process-wide "done" flag would skip the second device of a process)
    MLVAE_CHECK_CUDA(cudaFuncSetAttribute(linear_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    linear_fwd_kernel<<<(M + 127) / 128, kLinThreads, smem, (cudaStream_t)stream>>>(prm);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

}  // extern "C"
