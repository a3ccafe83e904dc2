// tcgen05 / TMEM dense projections for sm_100a.
//   mlvae_tc05_selftest : one 128 x N x K tile through the tensor-core path (descriptor / TMEM
//                         layout check used by tests/test_tc05_gpu.py)
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

}  // extern "C"
