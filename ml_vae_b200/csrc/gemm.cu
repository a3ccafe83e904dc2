// TMA-fed tcgen05 GEMM for the time-parallel matrix products of the training step (sm_100a):
//   LSTM input projection   P  = x W_ih^T + b          (modules/decoder.py:14-15,22 -> nn.LSTM's x W_ih^T for all t)
//   LSTM input gradient     dx = dA W_ih               [+ the inter-layer dropout mask of the layer below, fused]
//   LSTM weight gradients   dW_ih = dA^T x,  dW_hh = dA^T h_prev   (float32, ACCUMULATED straight into the parameters'
//                           gradient tensors in torch's gate order: no unpack / add kernels)
//   dense weight gradients  dW = g^T x   (modules/fc_block.py:9-16), split-K with a deterministic second pass
//   wide Linear forward     y = act(x W^T + b) for K >= 512
// These were cuBLAS calls through torch in round 1 (~1.0 ms of the 5.9 ms step).
//
// One persistent, warp-specialised kernel (320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor (3-D tensor maps, SWIZZLE_128B) into a ring of kStages shared-memory
//               stages {A 128 x 64, B BN x 64}, completion counted in bytes on the stage's "full" mbarrier
//   warp 1      MMA issuer: waits "full", issues 4 x tcgen05.mma (K = 16 each) per stage from uniform registers, commits to
//               the stage's "empty" mbarrier (the stage is refilled when the tensor core has read it) and, after the last
//               K step of a tile, to the accumulator's "full" mbarrier
//   warps 2..9  epilogue: tcgen05.ld the 128 x BN float32 accumulator (one of TWO TMEM buffers, so the mainloop of the next
//               tile overlaps the epilogue of this one), bias / LeakyReLU / dropout mask / dtype / row permutation; bf16
//               outputs leave through swizzled shared-memory staging + TMA stores (coalesced, clipped at the matrix edges),
//               float32 outputs (weight gradients, small) with direct 16-byte stores
// Operands may be K-major (row-major [rows][K]) or MN-major (row-major [K][rows], e.g. dA^T without a transpose pass): the
// major-ness goes into the instruction descriptor and the shared-memory descriptors (tc05.cuh).  MN-major operands take a
// batched reduction dimension (K rows per batch, `kbatches` batches): TMA zero-fills rows past the end of a batch, which is
// what makes dW_hh = sum_b sum_{t<T-1} dA[b,t+1]^T h[b,t] ONE GEMM over row-shifted views without boundary corrections.
#include <cuda.h>

#include <cstring>

#include "common.cuh"
#include "philox.cuh"
#include "tc05.cuh"

namespace mlvae {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kEpiWarps = 8;                 // two per TMEM lane quarter, each takes half of the accumulator columns
constexpr int kGemmThreads = (2 + kEpiWarps) * 32;
constexpr int kBM = 128, kBK = 64;
constexpr int kMaxProb = 4;
constexpr int kMaxStages = 8;

struct GemmProblem {
    CUtensorMap ta, tb, td;          // td: the bf16 output (TMA store), unused for float32 outputs
    void *D;
    const float *bias;
};
struct GemmParams {
    GemmProblem prob[kMaxProb];
    float *ws;                       // split-K partials [split][prob][M][N] float32
    const uint64_t *drop_offset_add;
    uint64_t drop_seed, drop_offset;
    int64_t ldd;
    int nprob, M, N, K, kbatches, a_mn, b_mn;
    int out_f32, accumulate, leaky, row_perm_H, split_k;
    int tiles_m, tiles_n, stages;
    uint32_t drop_thresh;            // 0 = no dropout epilogue
    float drop_scale;
    int debug_mode;                  // probes only: 1 = epilogue skips the global stores, 2 = skips the bias
};

struct WorkItem {
    int prob, split, m_blk, n_blk, it0, it1;
};
__device__ __forceinline__ WorkItem decode_work(const GemmParams &p, int w, int total_iters) {
    WorkItem x;
    x.n_blk = w % p.tiles_n;
    int r = w / p.tiles_n;
    x.m_blk = r % p.tiles_m;
    r /= p.tiles_m;
    x.split = r % p.split_k;
    x.prob = r / p.split_k;
    const int per = (total_iters + p.split_k - 1) / p.split_k;
    x.it0 = x.split * per;
    x.it1 = min(total_iters, x.it0 + per);
    return x;
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_bf16_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t s_full[kMaxStages], s_empty[kMaxStages], s_acc_full[2], s_acc_empty[2];
    __shared__ uint32_t s_tmem;
    // dynamic shared memory is only guaranteed 16-byte aligned: round up to the 1024-byte swizzle atom
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int kABytes = kBM * kBK * 2, kBBytes = BN * kBK * 2, kStageBytes = kABytes + kBBytes;
    constexpr uint32_t kTmemCols = (2 * BN) < 32 ? 32 : 2 * BN;
    __shared__ __align__(16) float s_bias[2][BN];           // the tile's bias slice, per accumulator buffer
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int S = p.stages;
    // bf16 outputs leave through shared memory + TMA stores: per epilogue warp one 32 row x 64 byte (32 column) staging
    // buffer behind the operand ring (SWIZZLE_64B, matching the output tensor map)
    unsigned char *s_out = smem + (size_t)S * kStageBytes;

    if (warp == 1) tc::tmem_alloc(&s_tmem, kTmemCols);
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            tc::mbar_init(&s_full[s], 1);
            tc::mbar_init(&s_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&s_acc_full[b], 1);
            tc::mbar_init(&s_acc_empty[b], kEpiWarps);
        }
        tc::fence_barrier_init();
    }
    if (warp == 0 && lane == 0)
        for (int i = 0; i < p.nprob; ++i) {
            tc::tma_prefetch_desc(&p.prob[i].ta);
            tc::tma_prefetch_desc(&p.prob[i].tb);
            if (!p.out_f32) tc::tma_prefetch_desc(&p.prob[i].td);
        }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0);

    const int kpb = (p.K + kBK - 1) / kBK;                 // K iterations per batch
    const int total_iters = p.kbatches * kpb;
    const int total_work = p.nprob * p.split_k * p.tiles_m * p.tiles_n;

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const WorkItem x = decode_work(p, w, total_iters);
            const CUtensorMap *ta = &p.prob[x.prob].ta, *tb = &p.prob[x.prob].tb;
            for (int it = x.it0; it < x.it1; ++it) {
                tc::mbar_wait(&s_empty[stage], phase ^ 1);
                if (tc::elect_one()) {
                    unsigned char *sa = smem + (size_t)stage * kStageBytes, *sb = sa + kABytes;
                    tc::mbar_arrive_expect_tx(&s_full[stage], kStageBytes);
                    const int kb = it / kpb, k0 = (it - kb * kpb) * kBK;
                    if (p.a_mn) {
#pragma unroll
                        for (int j = 0; j < kBM / 64; ++j) tc::tma_load_3d(sa + j * 8192, ta, &s_full[stage], x.m_blk * kBM + j * 64, k0, kb);
                    } else {
                        tc::tma_load_3d(sa, ta, &s_full[stage], k0, x.m_blk * kBM, 0);
                    }
                    if (p.b_mn) {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tc::tma_load_3d(sb + j * 8192, tb, &s_full[stage], x.n_blk * BN + j * 64, k0, kb);
                    } else {
                        tc::tma_load_3d(sb, tb, &s_full[stage], k0, x.n_blk * BN, 0);
                    }
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = tc::idesc_bf16_f32_major(kBM, BN, p.a_mn, p.b_mn);
        int stage = 0, buf = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const WorkItem x = decode_work(p, w, total_iters);
            tc::mbar_wait(&s_acc_empty[buf], acc_phase ^ 1);            // the epilogue has drained this accumulator buffer
            tc::fence_after_sync();
            const uint32_t d_tmem = tmem + buf * BN;
            for (int it = x.it0; it < x.it1; ++it) {
                tc::mbar_wait(&s_full[stage], phase);
                tc::fence_after_sync();
                if (tc::elect_one()) {
                    const uint32_t sa = tc::smem_u32(smem + (size_t)stage * kStageBytes), sb = sa + kABytes;
                    const uint64_t da = p.a_mn ? tc::smem_desc_sw128_mn(sa, 8192) : tc::smem_desc_sw128(sa);
                    const uint64_t db = p.b_mn ? tc::smem_desc_sw128_mn(sb, 8192) : tc::smem_desc_sw128(sb);
                    const uint64_t sta = p.a_mn ? 128 : 2, stb = p.b_mn ? 128 : 2;      // descriptor step per K = 16 (16-byte units)
#pragma unroll
                    for (int k16 = 0; k16 < kBK / 16; ++k16)
                        tc::mma_bf16(d_tmem, da + sta * k16, db + stb * k16, idesc, (it > x.it0) || (k16 > 0));
                    tc::mma_commit(&s_empty[stage]);                     // stage reusable once the tensor core has read it
                    if (it + 1 == x.it1) tc::mma_commit(&s_acc_full[buf]);
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
            if (x.it1 <= x.it0 && tc::elect_one()) tc::mma_commit(&s_acc_full[buf]);     // empty K range (never with valid arguments)
            buf ^= 1;
            if (buf == 0) acc_phase ^= 1;
        }
    } else {
        // ===================== epilogue (warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4) =====================
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int etid = tid - 64;                                          // 0..255 inside the epilogue warps
        int buf = 0;
        uint32_t acc_phase = 0;
        const PhiloxKey key(p.drop_seed);
        const uint64_t drop_off = p.drop_offset + (p.drop_offset_add ? *p.drop_offset_add : 0ull);
        unsigned char *my_out = s_out + (size_t)(warp - 2) * 2048;        // this warp's 32 row x 64 byte staging buffer
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const WorkItem x = decode_work(p, w, total_iters);
            const float *bias = (p.split_k > 1 || p.debug_mode == 2) ? nullptr : p.prob[x.prob].bias;
            if (bias) {                                                     // the tile's bias slice -> shared memory, once per tile
                for (int i = etid; i < BN; i += kEpiWarps * 32) {
                    const int n = x.n_blk * BN + i;
                    s_bias[buf][i] = n < p.N ? __ldg(bias + n) : 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");   // the two buffers alternate: no second barrier needed
            }
            tc::mbar_wait(&s_acc_full[buf], acc_phase);
            tc::fence_after_sync();
            const int m = x.m_blk * kBM + quarter * 32 + lane;
            const bool row_ok = m < p.M;
            const int out_row = p.row_perm_H > 0 ? (m & 3) * p.row_perm_H + (m >> 2) : m;
            constexpr int kHalves = 2;
            constexpr int kChunks = BN / 32 / kHalves;                      // 32-column chunks per warp
#pragma unroll 1
            for (int cc = 0; cc < (half < kHalves ? kChunks : 0); ++cc) {
                const int c = half * kChunks + cc;
                uint32_t v[32];
                tc::tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + buf * BN + c * 32, v);
                tc::tmem_ld_wait();
                const int n0 = x.n_blk * BN + c * 32;
                if (p.split_k > 1) {
                    if (row_ok && n0 < p.N) {
                        float *dst = p.ws + (((size_t)x.split * p.nprob + x.prob) * p.M + m) * p.N + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (n0 + j < p.N) *reinterpret_cast<float4 *>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    }
                    continue;
                }
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(&s_bias[buf][c * 32 + j]);     // broadcast read
                        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
                    }
                }
                if (p.leaky) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = f[j] > 0.f ? f[j] : 0.01f * f[j];
                }
                if (p.drop_thresh) {
                    // the keep mask of csrc/dropout.cu for element i = m * N + n (N % 8 == 0, ldd == N): one Philox block per 8 columns
                    const uint64_t i0 = (uint64_t)m * p.N + n0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint64_t qb = (i0 >> 3) + g;
                        const uint4 r = philox4x32_10(make_uint4((uint32_t)qb, (uint32_t)(qb >> 32), (uint32_t)drop_off, (uint32_t)(drop_off >> 32)), key);
                        const uint32_t wd[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const uint32_t u16 = (wd[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
                            f[g * 8 + e] = (u16 >= p.drop_thresh) ? f[g * 8 + e] * p.drop_scale : 0.f;
                        }
                    }
                }
                if (p.debug_mode == 1) continue;
                if (p.out_f32) {
                    if (!row_ok || n0 >= p.N) continue;
                    float *dst = reinterpret_cast<float *>(p.prob[x.prob].D) + (size_t)out_row * p.ldd + n0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (n0 + j < p.N) {
                            float4 o = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                            if (p.accumulate) {
                                const float4 old = *reinterpret_cast<const float4 *>(dst + j);
                                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                            }
                            *reinterpret_cast<float4 *>(dst + j) = o;
                        }
                    }
                } else {
                    // bf16: the 32 columns = 64 bytes of this lane's row go to the warp's staging buffer (32 rows x 64 bytes,
                    // SWIZZLE_64B like the output tensor map: 16-byte chunk ^= (row / 2) % 4, conflict-free) and the 32 x 32 block
                    // leaves as ONE TMA store (coalesced, clipped at the matrix edges by the tensor map)
                    if (lane == 0) tc::tma_store_wait_read<0>();            // the previous store has read the buffer
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[j + 2 * e], f[j + 2 * e + 1]);
                            pk[e] = *reinterpret_cast<const uint32_t *>(&h2);
                        }
                        *reinterpret_cast<uint4 *>(my_out + lane * 64 + (((j >> 3) ^ ((lane >> 1) & 3)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tc::tma_store_3d(&p.prob[x.prob].td, my_out, n0, x.m_blk * kBM + quarter * 32, 0);
                        tc::tma_store_commit();
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s_acc_empty[buf]);
            buf ^= 1;
            if (buf == 0) acc_phase ^= 1;
        }
        if (lane == 0) tc::tma_store_wait<0>();                             // all output stores complete before the CTA exits
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem, kTmemCols);
}

// second pass of split-K: D (+)= sum over the splits, in split order (deterministic)
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(const float *__restrict__ ws, int split_k, int nprob, int M, int N, int64_t ldd,
                                                                 int accumulate, int row_perm_H, GemmProblem p0, GemmProblem p1, GemmProblem p2,
                                                                 GemmProblem p3) {
    const int64_t per = (int64_t)M * (N / 4);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < per * nprob; i += (int64_t)gridDim.x * 256) {
        const int prob = (int)(i / per);
        const int64_t r = i - prob * per;
        const int m = (int)(r / (N / 4)), n = (int)(r - (int64_t)m * (N / 4)) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < split_k; ++s) {
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(ws + (((size_t)s * nprob + prob) * M + m) * N + n));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float *D = reinterpret_cast<float *>(prob == 0 ? p0.D : prob == 1 ? p1.D : prob == 2 ? p2.D : p3.D);
        const int out_row = row_perm_H > 0 ? (m & 3) * row_perm_H + (m >> 2) : m;
        float4 *dst = reinterpret_cast<float4 *>(D + (size_t)out_row * ldd + n);
        if (accumulate) {
            const float4 old = *dst;
            acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
        }
        *dst = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 3-D bf16 tensor map: dims (d0 innermost) with byte strides s1, s2; box (b0, b1, 1); SWIZZLE_128B; OOB reads give zeros
int make_map(CUtensorMap *map, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2, uint32_t b0, uint32_t b1,
             CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    MLVAE_REQUIRE(fn != nullptr, MLVAE_ERR_CUDA, "gemm: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {s1, s2};
    const cuuint32_t box[3] = {b0, b1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MLVAE_REQUIRE(r == CUDA_SUCCESS, MLVAE_ERR_CUDA, "gemm: cuTensorMapEncodeTiled failed with %d (dims %llu %llu %llu, strides %llu %llu, box %u %u)",
                  (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1, (unsigned long long)s2, b0, b1);
    return MLVAE_OK;
}

constexpr size_t kOutStageBytes = kEpiWarps * 2048;    // one 32 row x 64 byte staging buffer per epilogue warp
template <int BN>
int launch_gemm(const GemmParams &prm, int grid, cudaStream_t st) {
    const size_t stage = (size_t)kBM * kBK * 2 + (size_t)BN * kBK * 2;
    const size_t smem = (size_t)prm.stages * stage + (prm.out_f32 ? 0 : kOutStageBytes) + 1024;
    MLVAE_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_bf16_kernel<BN><<<grid, kGemmThreads, smem, st>>>(prm);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

namespace { int g_gemm_debug_mode = 0; }
extern "C" {

int mlvae_gemm_debug_mode(int mode) { g_gemm_debug_mode = mode; return MLVAE_OK; }

size_t mlvae_gemm_workspace_bytes(int nprob, int M, int N, int split_k) {
    return split_k > 1 ? (size_t)split_k * nprob * M * N * sizeof(float) : 0;
}

int mlvae_gemm_bf16(const mlvae_gemm_args *a, void *stream) {
    MLVAE_REQUIRE(a != nullptr, MLVAE_ERR_INVALID_ARG, "gemm: null arguments");
    MLVAE_REQUIRE(a->nprob >= 1 && a->nprob <= kMaxProb, MLVAE_ERR_INVALID_ARG, "gemm: nprob must be in [1, %d]", kMaxProb);
    MLVAE_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->kbatches >= 1, MLVAE_ERR_INVALID_ARG, "gemm: bad sizes");
    MLVAE_REQUIRE(a->N % 8 == 0 && a->lda % 8 == 0 && a->ldb % 8 == 0, MLVAE_ERR_UNSUPPORTED, "gemm: N, lda, ldb must be multiples of 8 (got %d %lld %lld)",
                  a->N, (long long)a->lda, (long long)a->ldb);
    MLVAE_REQUIRE(a->kbatches == 1 || (a->a_mn_major && a->b_mn_major), MLVAE_ERR_UNSUPPORTED, "gemm: a batched reduction needs both operands MN-major");
    MLVAE_REQUIRE(a->kbatches == 1 || (a->a_batch_stride % 8 == 0 && a->b_batch_stride % 8 == 0), MLVAE_ERR_UNSUPPORTED, "gemm: batch strides must be multiples of 8");
    MLVAE_REQUIRE(!a->a_mn_major || a->M % 8 == 0, MLVAE_ERR_UNSUPPORTED, "gemm: MN-major A needs M %% 8 == 0");
    MLVAE_REQUIRE(a->a_mn_major || a->K % 8 == 0, MLVAE_ERR_UNSUPPORTED, "gemm: K-major operands need K %% 8 == 0");
    MLVAE_REQUIRE(a->out_f32 || !a->accumulate, MLVAE_ERR_UNSUPPORTED, "gemm: accumulate needs float32 output");
    MLVAE_REQUIRE(a->row_perm_H == 0 || a->M == 4 * a->row_perm_H, MLVAE_ERR_INVALID_ARG, "gemm: the gate row permutation needs M == 4 H");
    int split = a->split_k > 1 ? a->split_k : 1;
    {
        const int total_iters = a->kbatches * ((a->K + kBK - 1) / kBK);
        if (split > total_iters) split = total_iters;
        const int per = (total_iters + split - 1) / split;
        split = (total_iters + per - 1) / per;                      // no empty split (the kernel cuts the range in chunks of `per`)
    }
    MLVAE_REQUIRE(split == 1 || (a->out_f32 && a->ws && a->N % 4 == 0 && a->drop_p == 0.f && !a->leaky), MLVAE_ERR_UNSUPPORTED,
                  "gemm: split-K writes float32 (+ workspace) and has no activation epilogue");
    MLVAE_REQUIRE(a->drop_p == 0.f || (a->ldd == a->N && a->drop_p > 0.f && a->drop_p < 1.f), MLVAE_ERR_UNSUPPORTED, "gemm: the dropout epilogue needs a dense output (ldd == N)");
    MLVAE_REQUIRE(a->ldd >= a->N && a->ldd % (a->out_f32 ? 4 : 8) == 0, MLVAE_ERR_INVALID_ARG, "gemm: bad ldd");

    int bn = a->bn;
    if (bn == 0) bn = a->N > 128 ? 256 : a->N > 64 ? 128 : 64;
    MLVAE_REQUIRE(bn == 64 || bn == 128 || bn == 256, MLVAE_ERR_INVALID_ARG, "gemm: bn must be 64, 128 or 256");

    GemmParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.nprob = a->nprob; prm.M = a->M; prm.N = a->N; prm.K = a->K; prm.kbatches = a->kbatches;
    prm.a_mn = a->a_mn_major ? 1 : 0; prm.b_mn = a->b_mn_major ? 1 : 0;
    prm.ldd = a->ldd; prm.out_f32 = a->out_f32; prm.accumulate = a->accumulate; prm.leaky = a->leaky; prm.row_perm_H = a->row_perm_H;
    prm.split_k = split; prm.ws = (float *)a->ws;
    prm.debug_mode = g_gemm_debug_mode;
    prm.tiles_m = (a->M + kBM - 1) / kBM; prm.tiles_n = (a->N + bn - 1) / bn;
    const size_t stage = (size_t)kBM * kBK * 2 + (size_t)bn * kBK * 2;
    int stages = (int)((220 * 1024 - (a->out_f32 ? 0 : kOutStageBytes)) / stage);
    prm.stages = stages > kMaxStages ? kMaxStages : stages;
    if (a->drop_p > 0.f) {
        prm.drop_thresh = (uint32_t)lrintf(a->drop_p * 65536.f);
        prm.drop_scale = 1.f / (1.f - a->drop_p);
        prm.drop_seed = a->drop_seed; prm.drop_offset = a->drop_offset; prm.drop_offset_add = (const uint64_t *)a->drop_offset_add;
    }
    for (int i = 0; i < a->nprob; ++i) {
        MLVAE_REQUIRE(a->A[i] && a->B[i] && a->D[i], MLVAE_ERR_INVALID_ARG, "gemm: null matrix (problem %d)", i);
        MLVAE_REQUIRE(((uintptr_t)a->A[i] & 15) == 0 && ((uintptr_t)a->B[i] & 15) == 0 && ((uintptr_t)a->D[i] & 15) == 0, MLVAE_ERR_INVALID_ARG,
                      "gemm: matrices must be 16-byte aligned");
        if (a->a_mn_major) {
            if (int rc = make_map(&prm.prob[i].ta, a->A[i], a->M, a->K, a->kbatches, (uint64_t)a->lda * 2,
                                  (uint64_t)(a->kbatches > 1 ? a->a_batch_stride : a->lda * (int64_t)a->K) * 2, 64, 64)) return rc;
        } else {
            if (int rc = make_map(&prm.prob[i].ta, a->A[i], a->K, a->M, 1, (uint64_t)a->lda * 2, (uint64_t)a->lda * 2 * a->M, 64, kBM)) return rc;
        }
        if (a->b_mn_major) {
            if (int rc = make_map(&prm.prob[i].tb, a->B[i], a->N, a->K, a->kbatches, (uint64_t)a->ldb * 2,
                                  (uint64_t)(a->kbatches > 1 ? a->b_batch_stride : a->ldb * (int64_t)a->K) * 2, 64, 64)) return rc;
        } else {
            if (int rc = make_map(&prm.prob[i].tb, a->B[i], a->K, a->N, 1, (uint64_t)a->ldb * 2, (uint64_t)a->ldb * 2 * a->N, 64, bn)) return rc;
        }
        if (!a->out_f32)
            if (int rc = make_map(&prm.prob[i].td, a->D[i], a->N, a->M, 1, (uint64_t)a->ldd * 2, (uint64_t)a->ldd * 2 * a->M, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
        prm.prob[i].D = a->D[i];
        prm.prob[i].bias = a->bias[i];
    }
    const int total = prm.nprob * split * prm.tiles_m * prm.tiles_n;
    const int sms = sm_count();
    int grid = total < sms ? total : sms;
    if (a->max_ctas > 0 && a->max_ctas < grid) grid = a->max_ctas;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = bn == 256 ? launch_gemm<256>(prm, grid, st) : bn == 128 ? launch_gemm<128>(prm, grid, st) : launch_gemm<64>(prm, grid, st);
    if (rc) return rc;
    if (split > 1) {
        const int64_t n4 = (int64_t)a->nprob * a->M * (a->N / 4);
        int64_t blocks = (n4 + 255) / 256;
        if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
        gemm_splitk_reduce_kernel<<<(int)blocks, 256, 0, st>>>(prm.ws, split, a->nprob, a->M, a->N, a->ldd, a->accumulate, a->row_perm_H, prm.prob[0], prm.prob[1],
                                                                prm.prob[2], prm.prob[3]);
        MLVAE_CHECK_CUDA(cudaGetLastError());
    }
    return MLVAE_OK;
}

}  // extern "C"
