// Fused two-layer Linear / LeakyReLU chains of the latent block:
//   modules/fc_block.py:4-21       FCBlock = Linear -> LeakyReLU -> Linear [-> LeakyReLU]
//   modules/vanilla_vae.py:13-24   encoder trunk  D -> 64 -> 64  (both activated)
//   modules/decoder.py:16-17,24-25 the tails of the two heads  64 -> 64 -> D  (last layer bare), both heads in ONE launch
// Round 1 ran every Linear as its own launch with an HBM round trip of the activation, and its backward as four launches
// (activation'/bias kernel, dx GEMM, dW GEMM, split-K reduce): 14 + 30 launches per step for layers whose whole weight set is
// a few KB.  Here a 128-row tile of the batch goes through BOTH layers on chip:
//   forward   x tile (TMA) -> tcgen05.mma vs W_A (smem) -> TMEM -> +bias, LeakyReLU, bf16 -> smem operand tile (+ HBM, saved for
//             backward) -> tcgen05.mma vs W_B -> TMEM -> +bias [, LeakyReLU] -> HBM
//   backward  g_out, y_A, x tiles (TMA) -> [g_B = g_out * act'(y_B)] -> dW_B|db_B += g_B^T [y_A | 1]  and  dy_A = g_B W_B
//             -> g_A = dy_A * LeakyReLU'(y_A) (smem operand tile) -> dW_A|db_A += g_A^T [x | 1]  and  dx = g_A W_A -> HBM
//   The weight / bias gradients accumulate IN TENSOR MEMORY over all tiles a CTA walks (the bias gradient is the extra
//   column that a ones-column appended to the activation tile produces), leave once per CTA as float32 partials and are
//   added in CTA order by a second kernel (deterministic, no atomics).  One shared-memory tile serves as K-major operand
//   (g W) and as MN-major operand (g^T y): a row-major [rows][64] SWIZZLE_128B tile is both.
// One CTA = 128 threads (thread = tile row = TMEM lane); the phases of a tile are sequential by construction, so the kernel
// is phase-synchronous (mbarrier for TMA / MMA completion, __syncthreads between phases) instead of warp-specialised.
#include <cuda.h>

#include <cstring>

#include "common.cuh"
#include "tc05.cuh"

namespace mlvae {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kChainThreads = 128;
constexpr int kTileRows = 128;
constexpr int kBlk = kTileRows * 128;            // bytes of one 64-column block of a 128-row tile
constexpr float kSlope = 0.01f;

// byte offset of element (r, c) of a row-major tile stored as 64-column SWIZZLE_128B blocks of `blk` bytes
__device__ __forceinline__ uint32_t tile_off(int r, int c, int blk = kBlk) {
    return (uint32_t)((c >> 6) * blk) + tc::sw128_off(r, (c & 63) >> 3) + (uint32_t)((c & 7) * 2);
}
// K-major view (rows = M / N index, columns = K): descriptor of the K = 16 step k16
__device__ __forceinline__ uint64_t desc_k(uint32_t base, int k16, int blk = kBlk) {
    return tc::smem_desc_sw128(base + (uint32_t)((k16 >> 2) * blk) + (uint32_t)((k16 & 3) * 32));
}
// MN-major view (columns = M / N index in 64-wide blocks `blk` apart, rows = K): descriptor of the K = 16 step k16
__device__ __forceinline__ uint64_t desc_mn(uint32_t base, int k16, int blk = kBlk) { return tc::smem_desc_sw128_mn(base + (uint32_t)(k16 * 2048), (uint32_t)blk); }

struct ChainProblemFwd {
    CUtensorMap tx, twa, twb;
    const float *bias_a, *bias_b;
    bf16 *ya, *yb;
};
struct ChainFwdParams {
    ChainProblemFwd prob[2];
    int M, KA, NA, NB, act_b, tiles;
    int64_t ld_ya, ld_yb;
};

__global__ void __launch_bounds__(kChainThreads, 1) chain2_fwd_kernel(const __grid_constant__ ChainFwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t s_tma, s_mma, s_w;
    __shared__ uint32_t s_tmem;
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sX = smem, *sY = smem + 2 * kBlk, *sWA = smem + 4 * kBlk, *sWB = smem + 6 * kBlk;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const ChainProblemFwd &q = p.prob[blockIdx.y];
    const int ka_blocks = (p.KA + 63) >> 6, na_blocks = (p.NA + 63) >> 6;

    if (warp == 0) tc::tmem_alloc(&s_tmem, 256);
    if (tid == 0) {
        tc::mbar_init(&s_tma, 1);
        tc::mbar_init(&s_mma, 1);
        tc::mbar_init(&s_w, 1);
        tc::fence_barrier_init();
        // weights: W_A (NA x KA) and W_B (NB x NA) as K-major operands, loaded once
        tc::mbar_arrive_expect_tx(&s_w, (uint32_t)(ka_blocks * p.NA * 128 + na_blocks * p.NB * 128));
        for (int j = 0; j < ka_blocks; ++j) tc::tma_load_3d(sWA + j * kBlk, &q.twa, &s_w, j * 64, 0, 0);
        for (int j = 0; j < na_blocks; ++j) tc::tma_load_3d(sWB + j * kBlk, &q.twb, &s_w, j * 64, 0, 0);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0);
    const uint32_t acc_a = tmem, acc_b = tmem + 128;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    tc::mbar_wait(&s_w, 0);
    uint32_t ph_tma = 0, ph_mma = 0;

    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int m0 = t * kTileRows;
        if (tid == 0) {
            tc::mbar_arrive_expect_tx(&s_tma, (uint32_t)(ka_blocks * kBlk));
            for (int j = 0; j < ka_blocks; ++j) tc::tma_load_3d(sX + j * kBlk, &q.tx, &s_tma, j * 64, m0, 0);
        }
        tc::mbar_wait(&s_tma, ph_tma);
        ph_tma ^= 1;
        // ---- layer A: acc_a = x W_A^T ----
        if (warp == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const uint32_t idesc = tc::idesc_bf16_f32(128, p.NA);
                const int ks = (p.KA + 15) >> 4;
                for (int k = 0; k < ks; ++k) tc::mma_bf16(acc_a, desc_k(tc::smem_u32(sX), k), desc_k(tc::smem_u32(sWA), k), idesc, k > 0);
                tc::mma_commit(&s_mma);
            }
            __syncwarp();
        }
        tc::mbar_wait(&s_mma, ph_mma);
        ph_mma ^= 1;
        tc::fence_after_sync();
        const int r = warp * 32 + lane, m = m0 + r;
        for (int c = 0; c < p.NA; c += 16) {
            uint32_t v[16];
            tc::tmem_ld16(lane_base + acc_a + c, v);
            tc::tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float a0 = __uint_as_float(v[2 * e]) + __ldg(q.bias_a + c + 2 * e), a1 = __uint_as_float(v[2 * e + 1]) + __ldg(q.bias_a + c + 2 * e + 1);
                a0 = a0 > 0.f ? a0 : kSlope * a0;
                a1 = a1 > 0.f ? a1 : kSlope * a1;
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
                pk[e] = *reinterpret_cast<const uint32_t *>(&h2);
            }
            *reinterpret_cast<uint4 *>(sY + tile_off(r, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4 *>(sY + tile_off(r, c + 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            if (m < p.M && q.ya) {
                *reinterpret_cast<uint4 *>(q.ya + (size_t)m * p.ld_ya + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4 *>(q.ya + (size_t)m * p.ld_ya + c + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
        }
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        // ---- layer B: acc_b = y_A W_B^T ----
        if (warp == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const uint32_t idesc = tc::idesc_bf16_f32(128, p.NB);
                for (int k = 0; k < (p.NA >> 4); ++k) tc::mma_bf16(acc_b, desc_k(tc::smem_u32(sY), k), desc_k(tc::smem_u32(sWB), k), idesc, k > 0);
                tc::mma_commit(&s_mma);
            }
            __syncwarp();
        }
        tc::mbar_wait(&s_mma, ph_mma);
        ph_mma ^= 1;
        tc::fence_after_sync();
        for (int c = 0; c < p.NB; c += 16) {
            uint32_t v[16];
            tc::tmem_ld16(lane_base + acc_b + c, v);
            tc::tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float a0 = __uint_as_float(v[2 * e]) + __ldg(q.bias_b + c + 2 * e), a1 = __uint_as_float(v[2 * e + 1]) + __ldg(q.bias_b + c + 2 * e + 1);
                if (p.act_b) {
                    a0 = a0 > 0.f ? a0 : kSlope * a0;
                    a1 = a1 > 0.f ? a1 : kSlope * a1;
                }
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
                pk[e] = *reinterpret_cast<const uint32_t *>(&h2);
            }
            if (m < p.M) {
                *reinterpret_cast<uint4 *>(q.yb + (size_t)m * p.ld_yb + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4 *>(q.yb + (size_t)m * p.ld_yb + c + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
        }
        tc::fence_before_sync();
        __syncthreads();                                         // the x / y tiles and both accumulators are free again
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// ======================================================================================
struct ChainProblemBwd {
    CUtensorMap tg, tya, tx, twb, twa;      // g_out (M x NB), y_A (M x NA), x (M x KA), W_B (NB x NA), W_A (NA x KA)
    const bf16 *yb;                          // (M x NB) output of layer B, only when that layer is activated
    bf16 *dx;                                // (M x KA) or nullptr
};
struct ChainBwdParams {
    ChainProblemBwd prob[2];
    float *ws;                               // [prob][cta][128][WS] float32 partials, WS = (NA + 16) + (KA + 16)
    int M, KA, NA, NB, act_b, tiles, ctas;
    int64_t ld_yb, ld_dx;
};

__global__ void __launch_bounds__(kChainThreads, 1) chain2_bwd_kernel(const __grid_constant__ ChainBwdParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t s_tma, s_mma, s_w;
    __shared__ uint32_t s_tmem;
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // 128-row tiles, two 64-column blocks each: g_B, y_A (+ ones column), x (+ ones column), g_A; then the weights
    unsigned char *sG = smem, *sYA = smem + 2 * kBlk, *sX = smem + 4 * kBlk, *sGA = smem + 6 * kBlk, *sWB = smem + 8 * kBlk, *sWA = smem + 10 * kBlk;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const ChainProblemBwd &q = p.prob[blockIdx.y];
    const int ka_blocks = (p.KA + 63) >> 6, na_blocks = (p.NA + 63) >> 6;
    const int wb_blk = p.NB * 128, wa_blk = p.NA * 128;          // weight tiles: NB (NA) rows of 128 bytes per 64-column block

    if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
    if (tid == 0) {
        tc::mbar_init(&s_tma, 1);
        tc::mbar_init(&s_mma, 1);
        tc::mbar_init(&s_w, 1);
        tc::fence_barrier_init();
        // W_B (NB rows x NA cols) and W_A (NA rows x KA cols), row-major: MN-major B operands of dy_A = g_B W_B and dx = g_A W_A
        tc::mbar_arrive_expect_tx(&s_w, (uint32_t)(na_blocks * wb_blk + ka_blocks * wa_blk));
        for (int j = 0; j < na_blocks; ++j) tc::tma_load_3d(sWB + j * wb_blk, &q.twb, &s_w, j * 64, 0, 0);
        for (int j = 0; j < ka_blocks; ++j) tc::tma_load_3d(sWA + j * wa_blk, &q.twa, &s_w, j * 64, 0, 0);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0);
    // accumulators (columns): dW_B|db_B [NB x (NA+16)], dW_A|db_A [NA x (KA+16)] persist over the tiles; dy_A, dx per tile
    const uint32_t acc_wb = tmem, acc_wa = tmem + 144, acc_dy = tmem + 288, acc_dx = tmem + 288 + 112;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    tc::mbar_wait(&s_w, 0);
    uint32_t ph_tma = 0, ph_mma = 0;
    bool first = true;
    const int r = warp * 32 + lane;

    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
        const int m0 = t * kTileRows, m = m0 + r;
        if (tid == 0) {
            // every tile is loaded as TWO blocks: columns past the matrix width arrive as zeros (room for the ones column)
            tc::mbar_arrive_expect_tx(&s_tma, (uint32_t)(6 * kBlk));
            for (int j = 0; j < 2; ++j) {
                tc::tma_load_3d(sG + j * kBlk, &q.tg, &s_tma, j * 64, m0, 0);
                tc::tma_load_3d(sYA + j * kBlk, &q.tya, &s_tma, j * 64, m0, 0);
                tc::tma_load_3d(sX + j * kBlk, &q.tx, &s_tma, j * 64, m0, 0);
            }
        }
        tc::mbar_wait(&s_tma, ph_tma);
        ph_tma ^= 1;
        // ---- prologue: g_B = g_out * act'(y_B) in place (activated last layer only); the ones columns (rows of the matrix only) ----
        if (p.act_b && m < p.M) {
            for (int c = 0; c < p.NB; c += 8) {
                uint4 gv = *reinterpret_cast<const uint4 *>(sG + tile_off(r, c));
                const uint4 yv = __ldg(reinterpret_cast<const uint4 *>(q.yb + (size_t)m * p.ld_yb + c));
                uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
                const uint32_t yw[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float lo = __uint_as_float(gw[e] << 16), hi = __uint_as_float(gw[e] & 0xffff0000u);
                    if (!(__uint_as_float(yw[e] << 16) > 0.f)) lo *= kSlope;
                    if (!(__uint_as_float(yw[e] & 0xffff0000u) > 0.f)) hi *= kSlope;
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
                    gw[e] = *reinterpret_cast<const uint32_t *>(&h2);
                }
                *reinterpret_cast<uint4 *>(sG + tile_off(r, c)) = make_uint4(gw[0], gw[1], gw[2], gw[3]);
            }
        }
        {
            const bf16 one = __float2bfloat16_rn(m < p.M ? 1.f : 0.f);
            *reinterpret_cast<bf16 *>(sYA + tile_off(r, p.NA)) = one;
            *reinterpret_cast<bf16 *>(sX + tile_off(r, p.KA)) = one;
        }
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        // ---- dW_B|db_B += g_B^T [y_A | 1]  (both MN-major, K = the tile's 128 rows);  dy_A = g_B W_B ----
        if (warp == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const uint32_t id_w = tc::idesc_bf16_f32_major(128, p.NA + 16, 1, 1);
                for (int k = 0; k < 8; ++k) tc::mma_bf16(acc_wb, desc_mn(tc::smem_u32(sG), k), desc_mn(tc::smem_u32(sYA), k), id_w, !first || k > 0);
                const uint32_t id_d = tc::idesc_bf16_f32_major(128, p.NA, 0, 1);
                for (int k = 0; k < ((p.NB + 15) >> 4); ++k)
                    tc::mma_bf16(acc_dy, desc_k(tc::smem_u32(sG), k), desc_mn(tc::smem_u32(sWB), k, wb_blk), id_d, k > 0);
                tc::mma_commit(&s_mma);
            }
            __syncwarp();
        }
        tc::mbar_wait(&s_mma, ph_mma);
        ph_mma ^= 1;
        tc::fence_after_sync();
        // ---- g_A = dy_A * LeakyReLU'(y_A) -> operand tile (columns NA..127 of its first block stay zero from the first pass) ----
        for (int c = 0; c < p.NA; c += 16) {
            uint32_t v[16];
            tc::tmem_ld16(lane_base + acc_dy + c, v);
            tc::tmem_ld_wait();
            const uint4 y0 = *reinterpret_cast<const uint4 *>(sYA + tile_off(r, c)), y1 = *reinterpret_cast<const uint4 *>(sYA + tile_off(r, c + 8));
            const uint32_t yw[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
            uint32_t pk[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float lo = __uint_as_float(v[2 * e]), hi = __uint_as_float(v[2 * e + 1]);
                if (!(__uint_as_float(yw[e] << 16) > 0.f)) lo *= kSlope;
                if (!(__uint_as_float(yw[e] & 0xffff0000u) > 0.f)) hi *= kSlope;
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
                pk[e] = *reinterpret_cast<const uint32_t *>(&h2);
            }
            *reinterpret_cast<uint4 *>(sGA + tile_off(r, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4 *>(sGA + tile_off(r, c + 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        // ---- dW_A|db_A += g_A^T [x | 1];  dx = g_A W_A ----
        if (warp == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const uint32_t id_w = tc::idesc_bf16_f32_major(128, p.KA + 16, 1, 1);
                for (int k = 0; k < 8; ++k) tc::mma_bf16(acc_wa, desc_mn(tc::smem_u32(sGA), k), desc_mn(tc::smem_u32(sX), k), id_w, !first || k > 0);
                if (q.dx) {
                    const uint32_t id_d = tc::idesc_bf16_f32_major(128, p.KA, 0, 1);
                    for (int k = 0; k < (p.NA >> 4); ++k)
                        tc::mma_bf16(acc_dx, desc_k(tc::smem_u32(sGA), k), desc_mn(tc::smem_u32(sWA), k, wa_blk), id_d, k > 0);
                }
                tc::mma_commit(&s_mma);
            }
            __syncwarp();
        }
        first = false;
        tc::mbar_wait(&s_mma, ph_mma);
        ph_mma ^= 1;
        tc::fence_after_sync();
        if (q.dx) {
            for (int c = 0; c < p.KA; c += 16) {
                uint32_t v[16];
                tc::tmem_ld16(lane_base + acc_dx + c, v);
                tc::tmem_ld_wait();
                if (m < p.M) {
                    uint32_t pk[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
                        pk[e] = *reinterpret_cast<const uint32_t *>(&h2);
                    }
                    *reinterpret_cast<uint4 *>(q.dx + (size_t)m * p.ld_dx + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4 *>(q.dx + (size_t)m * p.ld_dx + c + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();
    }
    // ---- this CTA's partial weight / bias gradients: row n of dW_B|db_B and of dW_A|db_A (thread = row) ----
    {
        const int WS = (p.NA + 16) + (p.KA + 16);
        float *dst = p.ws + (((size_t)blockIdx.y * p.ctas + blockIdx.x) * 128 + r) * WS;
        const bool any = blockIdx.x < p.tiles;                 // a CTA without tiles never touched its accumulators
        for (int c = 0; c < p.NA + 16; c += 16) {
            uint32_t v[16];
            tc::tmem_ld16(lane_base + acc_wb + c, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4 *>(dst + c + e) = any ? make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int c = 0; c < p.KA + 16; c += 16) {
            uint32_t v[16];
            tc::tmem_ld16(lane_base + acc_wa + c, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4 *>(dst + p.NA + 16 + c + e) = any ? make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3])) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

struct ChainGradOut {
    float *dw_b, *db_b, *dw_a, *db_a;
};
// dW / db += sum over the CTAs' partials, in CTA order
__global__ void __launch_bounds__(256) chain2_bwd_reduce_kernel(const float *__restrict__ ws, int ctas, int KA, int NA, int NB, ChainGradOut o0, ChainGradOut o1) {
    const ChainGradOut o = blockIdx.y == 0 ? o0 : o1;
    const int WS = (NA + 16) + (KA + 16);
    const float *base = ws + (size_t)blockIdx.y * ctas * 128 * WS;
    const int nB = NB * (NA + 1), nA = NA * (KA + 1);
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nB + nA; i += gridDim.x * 256) {
        int row, col, off;
        float *dst;
        if (i < nB) {
            row = i / (NA + 1); col = i - row * (NA + 1); off = col;
            dst = col < NA ? o.dw_b + (size_t)row * NA + col : o.db_b + row;
        } else {
            const int k = i - nB;
            row = k / (KA + 1); col = k - row * (KA + 1); off = NA + 16 + col;
            dst = col < KA ? o.dw_a + (size_t)row * KA + col : o.db_a + row;
        }
        float acc = 0.f;
        for (int c = 0; c < ctas; ++c) acc += base[((size_t)c * 128 + row) * WS + off];
        *dst += acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn chain_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)ptr;
    }
    return fn;
}
// row-major bf16 matrix (rows x cols, row stride ld elements) -> map with box [64 columns x box_rows rows], SWIZZLE_128B, zero fill
int chain_map(CUtensorMap *map, const void *base, int rows, int cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = chain_encode_fn();
    MLVAE_REQUIRE(fn != nullptr, MLVAE_ERR_CUDA, "mlp_chain: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)rows};
    const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MLVAE_REQUIRE(r == CUDA_SUCCESS, MLVAE_ERR_CUDA, "mlp_chain: cuTensorMapEncodeTiled failed with %d (rows %d cols %d ld %lld box %d)", (int)r, rows, cols,
                  (long long)ld, box_rows);
    return MLVAE_OK;
}

int check_dims(int M, int KA, int NA, int NB, int nprob) {
    MLVAE_REQUIRE(nprob == 1 || nprob == 2, MLVAE_ERR_INVALID_ARG, "mlp_chain: 1 or 2 problems per launch");
    MLVAE_REQUIRE(M > 0 && KA > 0 && NA > 0 && NB > 0, MLVAE_ERR_INVALID_ARG, "mlp_chain: bad sizes");
    MLVAE_REQUIRE(KA % 16 == 0 && NA % 16 == 0 && NB % 16 == 0 && KA <= 112 && NA <= 112 && NB <= 128, MLVAE_ERR_UNSUPPORTED,
                  "mlp_chain: widths must be multiples of 16 with K_A, N_A <= 112 and N_B <= 128 (got %d -> %d -> %d)", KA, NA, NB);
    return MLVAE_OK;
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

extern "C" {

int mlvae_mlp_chain_fwd(const mlvae_chain_fwd_args *a, void *stream) {
    MLVAE_REQUIRE(a != nullptr, MLVAE_ERR_INVALID_ARG, "mlp_chain_fwd: null arguments");
    if (int rc = check_dims(a->M, a->K_A, a->N_A, a->N_B, a->nprob)) return rc;
    MLVAE_REQUIRE(a->ld_x % 8 == 0 && a->ld_ya % 8 == 0 && a->ld_yb % 8 == 0 && a->ld_x >= a->K_A && a->ld_ya >= a->N_A && a->ld_yb >= a->N_B, MLVAE_ERR_INVALID_ARG,
                  "mlp_chain_fwd: leading dimensions must be multiples of 8 and cover the rows");
    ChainFwdParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.M = a->M; prm.KA = a->K_A; prm.NA = a->N_A; prm.NB = a->N_B; prm.act_b = a->act_b; prm.tiles = (a->M + kTileRows - 1) / kTileRows;
    prm.ld_ya = a->ld_ya; prm.ld_yb = a->ld_yb;
    for (int i = 0; i < a->nprob; ++i) {
        MLVAE_REQUIRE(a->x[i] && a->w_a[i] && a->w_b[i] && a->bias_a[i] && a->bias_b[i] && a->y_b[i], MLVAE_ERR_INVALID_ARG, "mlp_chain_fwd: missing buffers (problem %d)", i);
        if (int rc = chain_map(&prm.prob[i].tx, a->x[i], a->M, a->K_A, a->ld_x, kTileRows)) return rc;
        if (int rc = chain_map(&prm.prob[i].twa, a->w_a[i], a->N_A, a->K_A, a->K_A, a->N_A)) return rc;
        if (int rc = chain_map(&prm.prob[i].twb, a->w_b[i], a->N_B, a->N_A, a->N_A, a->N_B)) return rc;
        prm.prob[i].bias_a = a->bias_a[i]; prm.prob[i].bias_b = a->bias_b[i];
        prm.prob[i].ya = (bf16 *)a->y_a[i]; prm.prob[i].yb = (bf16 *)a->y_b[i];
    }
    const int per = sm_count() / a->nprob;
    dim3 grid(prm.tiles < per ? prm.tiles : per, a->nprob);
    const size_t smem = 8 * (size_t)kBlk + 1024;
    MLVAE_CHECK_CUDA(cudaFuncSetAttribute(chain2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chain2_fwd_kernel<<<grid, kChainThreads, smem, (cudaStream_t)stream>>>(prm);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

size_t mlvae_mlp_chain_bwd_workspace_bytes(int nprob, int K_A, int N_A) {
    return (size_t)nprob * sm_count() * 128 * ((N_A + 16) + (K_A + 16)) * sizeof(float);
}

int mlvae_mlp_chain_bwd(const mlvae_chain_bwd_args *a, void *stream) {
    MLVAE_REQUIRE(a != nullptr, MLVAE_ERR_INVALID_ARG, "mlp_chain_bwd: null arguments");
    if (int rc = check_dims(a->M, a->K_A, a->N_A, a->N_B, a->nprob)) return rc;
    MLVAE_REQUIRE(a->ws != nullptr, MLVAE_ERR_INVALID_ARG, "mlp_chain_bwd: workspace missing");
    MLVAE_REQUIRE(a->ld_g % 8 == 0 && a->ld_ya % 8 == 0 && a->ld_x % 8 == 0 && a->ld_yb % 8 == 0 && a->ld_dx % 8 == 0, MLVAE_ERR_INVALID_ARG,
                  "mlp_chain_bwd: leading dimensions must be multiples of 8");
    ChainBwdParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.M = a->M; prm.KA = a->K_A; prm.NA = a->N_A; prm.NB = a->N_B; prm.act_b = a->act_b; prm.tiles = (a->M + kTileRows - 1) / kTileRows;
    prm.ld_yb = a->ld_yb; prm.ld_dx = a->ld_dx; prm.ws = (float *)a->ws;
    const int per = sm_count() / a->nprob;
    prm.ctas = prm.tiles < per ? prm.tiles : per;
    ChainGradOut outs[2];
    memset(outs, 0, sizeof(outs));
    for (int i = 0; i < a->nprob; ++i) {
        MLVAE_REQUIRE(a->g_out[i] && a->y_a[i] && a->x[i] && a->w_a[i] && a->w_b[i] && a->dw_a[i] && a->db_a[i] && a->dw_b[i] && a->db_b[i], MLVAE_ERR_INVALID_ARG,
                      "mlp_chain_bwd: missing buffers (problem %d)", i);
        MLVAE_REQUIRE(!a->act_b || a->y_b[i], MLVAE_ERR_INVALID_ARG, "mlp_chain_bwd: an activated last layer needs its output y_b");
        if (int rc = chain_map(&prm.prob[i].tg, a->g_out[i], a->M, a->N_B, a->ld_g, kTileRows)) return rc;
        if (int rc = chain_map(&prm.prob[i].tya, a->y_a[i], a->M, a->N_A, a->ld_ya, kTileRows)) return rc;
        if (int rc = chain_map(&prm.prob[i].tx, a->x[i], a->M, a->K_A, a->ld_x, kTileRows)) return rc;
        if (int rc = chain_map(&prm.prob[i].twb, a->w_b[i], a->N_B, a->N_A, a->N_A, a->N_B)) return rc;
        if (int rc = chain_map(&prm.prob[i].twa, a->w_a[i], a->N_A, a->K_A, a->K_A, a->N_A)) return rc;
        prm.prob[i].yb = (const bf16 *)a->y_b[i];
        prm.prob[i].dx = (bf16 *)a->dx[i];
        outs[i] = ChainGradOut{a->dw_b[i], a->db_b[i], a->dw_a[i], a->db_a[i]};
    }
    dim3 grid(prm.ctas, a->nprob);
    const size_t smem = 12 * (size_t)kBlk + 1024;
    MLVAE_CHECK_CUDA(cudaFuncSetAttribute(chain2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaStream_t st = (cudaStream_t)stream;
    chain2_bwd_kernel<<<grid, kChainThreads, smem, st>>>(prm);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    const int items = a->N_B * (a->N_A + 1) + a->N_A * (a->K_A + 1);
    chain2_bwd_reduce_kernel<<<dim3((items + 255) / 256, a->nprob), 256, 0, st>>>(prm.ws, prm.ctas, a->K_A, a->N_A, a->N_B, outs[0], outs[1]);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

}  // extern "C"
