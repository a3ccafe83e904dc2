// Persistent bidirectional LSTM recurrence for sm_100a (the decoder's
// 2 x biLSTM(512): modules/decoder.py:14-15,22).  cuDNN runs one GEMM + one cell kernel per
// timestep (2 x T launches per direction-layer); here ONE cooperative launch walks all T steps.
//
// Work split.  A "group" = (direction d, batch slice of NB rows).  Its G = H/32 CTAs each own 32
// hidden units = 128 gate rows (row = gate*32 + unit) of W_hh.  The slice (128 x H bf16) is loaded
// ONCE into TENSOR MEMORY (H/2 32-bit columns) and stays there for the whole sequence, so the
// per-step MMA
//     D[128 gate rows x NB batch] (TMEM, fp32) = W_slice (TMEM) x h_{t-1}^T (smem, K-major)
// does not stream 128 KB of weights through shared memory (v1 did: 2700 cycles/step; the TMEM-A
// form is bounded by 128*NB/256 cycles per K=16 instruction).  The input projection
// P = x W_ih^T + b_ih + b_hh for all timesteps is one big GEMM done before the launch.
//
// Exchange of h_t inside a group uses a flag-in-data protocol over L2 (no fence / atomic / separate
// flag round trip).  The step time is set by this exchange and it scales with the bytes every CTA
// pulls per step (tests/probes/exchange_probe.cu: 960 / 1330 / 2040 cycles for 8 / 16 / 32 KB), so
// the words carry NO separate tag: |h| <= 1 leaves bit 14 (the exponent MSB) of every bf16 free, and
// that bit of ALL eight elements of a 16-byte word holds the step tag ((step + 1) >> 1) & 1 -- it
// alternates between successive uses of a slot and differs from the zeroed initial state.  Every
// element is validated on its own, so the protocol does not even depend on 16-byte store atomicity.
// A NaN h (the only value with bit 14 set) travels as 0; the output Y keeps the NaN, so the loss is
// non-finite exactly when the reference's is.  Consumers write the words into the K-major B-operand
// tile and elected threads issue the tcgen05.mma instructions.  The backward pass ships its partial
// dh sums (unbounded, so no free bit) as 8-byte {2 x bf16, step tag} words; a 16-byte bit-tagged
// variant (values scaled by 2^-64) was measured 8 % slower there.  Deterministic; no data atomics.
// All CTAs wait on each other, so the launch is cooperative (co-residency guaranteed or refused).
#include <cooperative_groups.h>

#include "common.cuh"
#include "tc05.cuh"

namespace mlvae {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kUnits = 32;            // hidden units per CTA
constexpr int kRows = 4 * kUnits;     // gate rows per CTA == MMA M
constexpr int kLstmThreads = 512;
constexpr int kLstmWarps = kLstmThreads / 32;

#ifndef MLVAE_LSTM_EXACT_ACT
// one MUFU.TANH per activation (max relative error 2^-11, below the bf16 rounding of h and of the saved gates)
__device__ __forceinline__ float tanh_f(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_f(float x) { return fmaf(0.5f, tanh_f(0.5f * x), 0.5f); }
#else
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return __fdividef(2.f, 1.f + __expf(-2.f * x)) - 1.f; }
#endif

__device__ __forceinline__ uint2 ld_volatile_u2(const uint2 *p) {
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u2(uint2 *p, uint2 v) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_u4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u4(uint4 *p, uint4 v) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
constexpr uint32_t kTagBits = 0x40004000u;                       // bit 14 of both bf16 halves
__device__ __forceinline__ uint32_t step_tag(int step) { return (((step + 1) >> 1) & 1) ? kTagBits : 0u; }
__device__ __forceinline__ bool tag_ok(const uint4 &w, uint32_t tag) {
    return (((w.x & kTagBits) == tag) & ((w.y & kTagBits) == tag)) & (((w.z & kTagBits) == tag) & ((w.w & kTagBits) == tag));
}

__device__ long long *g_prof = nullptr;     // optional per-phase cycle counters (debug / profiles/)
#define PROF_MARK(k)                                                     \
    do {                                                                 \
        if (prof && tid == 0) { const long long now = clock64(); prof[k] += now - tprev; tprev = now; } \
    } while (0)

struct LstmFwdParams {
    bf16 *P;                 // (B, T, 2, H, 4) gate pre-activations (unit-major, the 4 gates i,f,g,o adjacent) from the input projection;
                             // overwritten with the ACTIVATED gates (i, f, g, o) when save != 0
    const bf16 *Whh;         // (2, 4H, H)
    bf16 *Y;                 // (B, T, 2H)
    float *C;                // (B, T, 2H) cell states (saved for backward) or nullptr
    uint4 *ll;               // [2 parity][2 * slices groups][G producers][NB rows][4] zeroed words of 8 tagged bf16
    int B, T, H, save;
};

template <int NB, int NCH>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_fwd_kernel(LstmFwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int H = p.H, T = p.T, B = p.B;
    const int u = blockIdx.x;                     // unit slice (gridDim.x == H / 32)
    const int slice = blockIdx.y;                 // batch slice
    const int d = blockIdx.z;                     // direction
    const int b0 = slice * NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3;                       // TMEM lane group of this warp
    const int part = warp >> 2;                   // which quarter of the accumulator columns (batch rows)
    constexpr int CPW = NB / 4;                   // batch rows per warp
    static_assert(CPW == 4, "the quad transpose below assumes 4 batch rows per warp (NB == 16)");
    // Gate rows are interleaved so that the 4 gates of a unit sit in 4 adjacent TMEM lanes:
    //   CTA row r = unit_local * 4 + gate  ->  lane = r % 32 of group q = r / 32
    // After the MMA a quad of lanes holds {i, f, g, o} x 4 batch rows of one unit; a 4 x 4 shuffle transpose
    // gives every lane all four gates of ONE (unit, batch row): no shared-memory round trip, no block barrier
    // between the non-linearity and the cell update.
    const int gate = lane & 3;
    const int unit_local = 8 * q + (lane >> 2);
    // NCH issuer warps, each feeding its own accumulator tile with the K steps k = w (mod NCH): issuing the
    // H/16 small MMAs from one thread is instruction-issue bound (descriptor arithmetic on the uniform
    // datapath), while more tiles cost TMEM read bandwidth in the epilogue (64 B/cycle).

    unsigned char *sH = smem;                                               // NB x H bf16, K-major

    const uint32_t dcol = (uint32_t)((H / 2 + 31) & ~31);                   // accumulator columns start
    uint32_t tmem_cols = 32;
    while (tmem_cols < dcol + NCH * NB) tmem_cols <<= 1;
    if (warp == 0) tc::tmem_alloc(&s_tmem, tmem_cols);
    if (tid == 0) {
        tc::mbar_init(&s_bar, NCH);
        tc::fence_barrier_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;

    // ---- resident weights: W_hh[d][gate*H + u*32 + unit_local][:] -> this thread's TMEM lane, columns k/2 ----
    {
        const bf16 *wrow = p.Whh + ((size_t)d * 4 * H + (size_t)gate * H + u * kUnits + unit_local) * H;
        for (int k16 = part; k16 < H / 16; k16 += 4) {
            const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(wrow + k16 * 16));
            const uint4 hi = __ldg(reinterpret_cast<const uint4 *>(wrow + k16 * 16 + 8));
            const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            tc::tmem_st8(tmem + lane_base + k16 * 8, v);
        }
        tc::tmem_st_wait();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    const uint32_t idesc = tc::idesc_bf16_f32(kRows, NB);
    const uint32_t sbo = (uint32_t)(H >> 3) * 128;
    const int groups = 2 * gridDim.y;
    const int group = d * gridDim.y + slice;
    const int G = H / kUnits;                                                 // CTAs (producers) per group
    const size_t ll_words = (size_t)G * NB * 4;                               // 16-byte words per group and parity

    // ---- per-thread constant addressing, hoisted out of the time loop ----
    const int t_first = d ? (T - 1) : 0;
    const ptrdiff_t p_step = (ptrdiff_t)(d ? -1 : 1) * 2 * H;               // P stride per time step in 8-byte units
    const ptrdiff_t y_step = (ptrdiff_t)(d ? -1 : 1) * 2 * H;
    const int my_row = part * CPW + gate;                                    // batch row this lane owns after the transpose
    const bool my_ok = (b0 + my_row) < B;
    const size_t y_off = ((size_t)min(b0 + my_row, B - 1) * T + t_first) * (2 * H) + d * H + u * kUnits + unit_local;
    bf16 *pY = p.Y + y_off;
    float *pC = p.C ? p.C + y_off : nullptr;
    // gate buffer layout: (B, T, 2, H, 4) -- the four gates of a unit are adjacent, so the lane that owns (unit, batch
    // row) after the transpose reads its pre-activations and writes its activated gates with ONE 8-byte access
    uint2 *pG = reinterpret_cast<uint2 *>(p.P) +
                (((size_t)min(b0 + my_row, B - 1) * T + t_first) * 2 + d) * (size_t)H + u * kUnits + unit_local;
    // this warp's eight units of batch row my_row are one exchange word: producer u, row my_row, quarter q, element lane >> 2
    const size_t ll_mine = ((size_t)u * NB + my_row) * 4 + q;
    const bool publisher = (lane >> 2) == 0;
    // consumer side: warp w pulls the NB x 4 words of producer w (1 KB, contiguous): word n * 32 + lane -> row, quarter
    constexpr int WB = NB * 4 / 32;
    uint32_t soff[WB];
#pragma unroll
    for (int n = 0; n < WB; ++n) {
        const int i = n * 32 + lane;
        soff[n] = tc::kmajor_off(i >> 2, 8 * (4 * (warp < G ? warp : 0) + (i & 3)), H);
    }
    float c_state = 0.f;

    // input-projection terms are prefetched one step ahead as RAW bits (converting at load time would stall the warp on
    // the DRAM latency inside the step)
    uint2 pre_raw = *pG;

    long long *prof = (g_prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_prof : nullptr;
    long long tprev = clock64();

    for (int step = 0; step < T; ++step) {
        float acc[CPW];
        if (step > 0) {
            // ---- gather h_{t_prev}: the {data, tag} words of this group, tag == step ----
            // Warp w pulls producer w's words (1 KB, coalesced).  Every thread spins on its FIRST word only and then fetches
            // the second: polling everything saturates L2 (128 CTAs x 16 KB per ~300-cycle round > the ~6 KB/cycle L2 cap)
            // and delays the producers' stores (1.00 ms), a few representative pollers per warp cost an extra round trip
            // whenever the other sectors land later (0.97 ms); this form measured 0.94 ms (8-byte {data, tag} words: 1.04 ms).
            if (warp < G) {
                const uint4 *src = p.ll + ((size_t)(step & 1) * groups + group) * ll_words + (size_t)warp * (NB * 4);
                const uint32_t tag = step_tag(step);
                uint4 w[WB];
                w[0] = ld_volatile_u4(src + lane);
                while (!tag_ok(w[0], tag)) w[0] = ld_volatile_u4(src + lane);
#pragma unroll
                for (int n = 1; n < WB; ++n) w[n] = ld_volatile_u4(src + n * 32 + lane);
#pragma unroll
                for (int n = 0; n < WB; ++n) {
                    while (!tag_ok(w[n], tag)) w[n] = ld_volatile_u4(src + n * 32 + lane);
                    *reinterpret_cast<uint4 *>(sH + soff[n]) =
                        make_uint4(w[n].x & ~kTagBits, w[n].y & ~kTagBits, w[n].z & ~kTagBits, w[n].w & ~kTagBits);
                }
            }
            tc::fence_proxy_async();
            tc::fence_before_sync();
            __syncthreads();
            PROF_MARK(0);                               // gather (includes waiting for the slowest producer)
            if (warp < NCH) {                            // warp-uniform: issuer warps
                if (tc::elect_one()) {
                    tc::fence_after_sync();
                    const uint64_t b_desc0 = tc::smem_desc(tc::smem_u32(sH), 128, sbo);
                    const uint32_t d_tile = tmem + dcol + warp * NB;
#pragma unroll 4
                    for (int k = warp; k < H / 16; k += NCH)
                        tc::mma_bf16_ts(d_tile, tmem + k * 8, b_desc0 + (uint64_t)(k * 16), idesc, k >= NCH);
                    tc::mma_commit(&s_bar);
                }
                __syncwarp();
            }
            tc::mbar_wait(&s_bar, (step - 1) & 1);
            tc::fence_after_sync();
            PROF_MARK(1);                               // MMA issue + completion
            uint32_t v[NCH][CPW];
#pragma unroll
            for (int c = 0; c < NCH; ++c) tc::tmem_ld<CPW>(tmem + lane_base + dcol + c * NB + part * CPW, v[c]);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < CPW; ++i) {
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    if (c < H / 16) a += __uint_as_float(v[c][i]);
                acc[i] = a;
            }
            tc::fence_before_sync();                     // orders these TMEM reads before the next step's MMAs
        } else {
#pragma unroll
            for (int i = 0; i < CPW; ++i) acc[i] = 0.f;      // h_0 = 0
        }
        // ---- 4 x 4 transpose of the raw accumulators inside each quad: lane `gate` ends up with the {i, f, g, o}
        //      pre-activations of batch row part*4 + gate for its unit ----
        float a[CPW];
#pragma unroll
        for (int i = 0; i < CPW; ++i) a[i] = acc[i];
        {
            const bool b0_ = lane & 1, b1_ = lane & 2;
            float x = b0_ ? a[0] : a[1], y = b0_ ? a[2] : a[3];
            x = __shfl_xor_sync(0xffffffffu, x, 1);
            y = __shfl_xor_sync(0xffffffffu, y, 1);
            if (b0_) { a[0] = x; a[2] = y; } else { a[1] = x; a[3] = y; }
            x = b1_ ? a[0] : a[2];
            y = b1_ ? a[1] : a[3];
            x = __shfl_xor_sync(0xffffffffu, x, 2);
            y = __shfl_xor_sync(0xffffffffu, y, 2);
            if (b1_) { a[0] = x; a[1] = y; } else { a[2] = x; a[3] = y; }
        }
        // ---- add the input projection (one 8-byte word: 4 gates) and apply the non-linearities ----
        a[0] = sigmoid_f(a[0] + __uint_as_float(pre_raw.x << 16));
        a[1] = sigmoid_f(a[1] + __uint_as_float(pre_raw.x & 0xffff0000u));
        a[2] = tanh_f(a[2] + __uint_as_float(pre_raw.y << 16));
        a[3] = sigmoid_f(a[3] + __uint_as_float(pre_raw.y & 0xffff0000u));
        if (step + 1 < T) pre_raw = *(pG + p_step);      // prefetch next step's input projection (a full step to land)
        // ---- cell update + publish ----
        const float c = a[1] * c_state + a[0] * a[2];
        c_state = c;
        const float h = a[3] * tanh_f(c);
        const bf16 hb = __float2bfloat16_rn(h);
        {
            uint32_t mine = (uint32_t)__bfloat16_as_ushort(hb);
            if (mine & 0x4000u) mine = 0;                                    // NaN (|h| <= 1 otherwise): travels as 0, Y keeps it
            mine |= step_tag(step + 1) & 0xffffu;
            // assemble the 8 units of this batch row (lanes lane&3 + 4e) in lane e == 0: e0|e1, e2|e3, e4|e5, e6|e7
            const uint32_t pair = mine | (__shfl_xor_sync(0xffffffffu, mine, 4) << 16);
            const uint32_t y2 = __shfl_xor_sync(0xffffffffu, pair, 8);
            const uint32_t z2 = __shfl_xor_sync(0xffffffffu, pair, 16);
            const uint32_t w2 = __shfl_xor_sync(0xffffffffu, y2, 16);
            if (publisher && step + 1 < T)
                st_volatile_u4(p.ll + ((size_t)((step + 1) & 1) * groups + group) * ll_words + ll_mine, make_uint4(pair, y2, z2, w2));
        }
        // ---- everything below is off the critical path: it overlaps the L2 flight time of the words just published ----
        if (my_ok) {
            *pY = hb;
            if (pC) *pC = c;
        }
        pY += y_step;
        if (pC) pC += y_step;
        if (p.save && my_ok) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(a[0], a[1]), hi = __floats2bfloat162_rn(a[2], a[3]);
            *pG = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
        }
        pG += p_step;
        PROF_MARK(2);                                   // activation + cell update + stores
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, tmem_cols);
}

// ======================================================================================
// Backward recurrence.  Same groups / ownership as the forward pass.  Per step (reverse order):
//   phase A  thread = (batch row j, unit): dh = dY + dh_rec; gate gradients -> pre-activation
//            gradients da_{i,f,g,o}; written (bf16) over the saved gates in G (input of the weight /
//            input GEMMs that follow) and into the K-major B-operand tile (NB x 128 gate rows)
//   phase B  partial[jh, j] = sum_{r in my 128 gate rows} W_hh[r, jh] * da[j, r]   tcgen05.mma with the
//            TRANSPOSED weight slice resident in TMEM (ceil(H/128) M-tiles x 64 columns)
//   phase C  partials leave as {2 x bf16, tag} words addressed to the CTA that owns unit jh; every
//            CTA sums the G partial slices it receives in fixed producer order (deterministic)
// ======================================================================================
struct LstmBwdParams {
    bf16 *G;                 // (B, T, 2, H, 4) in: activated gates from the forward pass; out: pre-activation grads
    const float *C;          // (B, T, 2H) cell states from the forward pass
    const bf16 *dY;          // (B, T, 2H) gradient of the layer output
    const bf16 *Whh;         // (2, 4H, H)
    uint2 *ll;               // [2 parity][groups][G consumers][G producers][NB/2][32] zeroed {data, tag} words
    float *db_part;          // (slices, 2, 4H) per-batch-slice sums over (rows, t) of dA, torch gate order; or nullptr
    int B, T, H;
};

template <int NB, int NI>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_bwd_kernel(LstmBwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    static_assert(NB == 16, "backward kernel is instantiated for NB = 16");
    const int H = p.H, T = p.T, B = p.B;
    const int G = H / kUnits;
    const int u = blockIdx.x, slice = blockIdx.y, d = blockIdx.z;
    const int b0 = slice * NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, part = warp >> 2;
    constexpr int CPW = NB / 4;
    const int tiles = (H + 127) / 128;
    const int issuers = tiles < NI ? tiles : NI;

    unsigned char *sDA = smem;                                           // NB x 128 bf16, K-major (4 KB)
    float2 *s_part = reinterpret_cast<float2 *>(smem + NB * 128 * 2);    // [2 halves][NB/2][32]

    const uint32_t dcol = (uint32_t)((tiles * 64 + 31) & ~31);
    uint32_t tmem_cols = 32;
    while (tmem_cols < dcol + (uint32_t)tiles * NB) tmem_cols <<= 1;
    if (warp == 0) tc::tmem_alloc(&s_tmem, tmem_cols);
    if (tid == 0) {
        tc::mbar_init(&s_bar, issuers);
        tc::fence_barrier_init();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;

    // ---- transposed weight slice -> TMEM: tile m, lane = jh - 128 m, K index r = gate*32 + unit ----
    {
        const unsigned short *W16 = reinterpret_cast<const unsigned short *>(p.Whh) + (size_t)d * 4 * H * H;
        for (int m = part; m < tiles; m += 4) {
            const int jh = 128 * m + 32 * q + lane;
            for (int k16 = 0; k16 < 8; ++k16) {
                uint32_t v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    uint32_t lo = 0, hi = 0;
                    if (jh < H) {
                        const int r0 = k16 * 16 + 2 * e, r1 = r0 + 1;
                        lo = W16[((size_t)(r0 >> 5) * H + u * kUnits + (r0 & 31)) * H + jh];
                        hi = W16[((size_t)(r1 >> 5) * H + u * kUnits + (r1 & 31)) * H + jh];
                    }
                    v[e] = lo | (hi << 16);
                }
                tc::tmem_st8(tmem + lane_base + m * 64 + k16 * 8, v);
            }
        }
        tc::tmem_st_wait();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    const uint32_t idesc = tc::idesc_bf16_f32(128, NB);
    const int groups = 2 * gridDim.y;
    const int group = d * gridDim.y + slice;
    const size_t ll_words = (size_t)G * G * (NB / 2) * 32;
    const int c_half = (G + 1) / 2;

    // phase-A identity of this thread
    const int j = warp, unit = lane;
    const int b = min(b0 + j, B - 1);                 // rows past B are clamped for loads, never stored
    const bool row_ok = (b0 + j) < B;
    const size_t g_row = (size_t)2 * 4 * H;
    uint2 *G2 = reinterpret_cast<uint2 *>(p.G);
    const unsigned short *dY16 = reinterpret_cast<const unsigned short *>(p.dY);
    // hoisted addressing: element offsets of (b, t, this unit) advance by a constant per step
    const int t_first = d ? 0 : (T - 1);
    const ptrdiff_t g_step = (ptrdiff_t)(d ? 1 : -1) * 2 * H;
    const ptrdiff_t y_step = (ptrdiff_t)(d ? 1 : -1) * 2 * H;
    size_t g_off = (((size_t)b * T + t_first) * 2 + d) * (size_t)H + u * kUnits + unit;        // 8-byte units: (B,T,2,H,4)
    size_t y_off = ((size_t)b * T + t_first) * (2 * H) + d * H + u * kUnits + unit;

    float dc_carry = 0.f;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};             // bias gradient: sum over t of this (row, unit)'s dA, per gate
    // raw prefetch for the first step
    int t = d ? 0 : (T - 1);
    uint2 rg;
    unsigned short rdy;
    float rc, rcp;
    {
        rg = G2[g_off];
        rdy = dY16[y_off];
        rc = p.C[y_off];
        rcp = (T > 1) ? p.C[y_off + y_step] : 0.f;          // c_{t-1} in forward order == next time index visited here
    }
    long long *prof = (g_prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_prof : nullptr;
    long long tprev = clock64();

    for (int step = 0; step < T; ++step) {
        t = d ? step : (T - 1 - step);
        float dh_rec = 0.f;
        if (step > 0) {
            // ---- phase C (consumer side): sum the partial slices addressed to this CTA ----
            const uint2 *src = p.ll + ((size_t)(step & 1) * groups + group) * ll_words + (size_t)u * G * (NB / 2) * 32;
            const int jp = warp & 7, half = warp >> 3;
            const int c_lo = half * c_half, c_hi = min(G, c_lo + c_half);
            float sx = 0.f, sy = 0.f;
            uint2 w[8];
            const uint2 *wsrc = src + ((size_t)c_lo * (NB / 2) + jp) * 32 + lane;       // producer stride: (NB/2)*32 words
            constexpr int PS = (NB / 2) * 32;
            // spin on one word, then fetch the rest (see the forward kernel)
            if (c_lo < c_hi) {
                w[0] = ld_volatile_u2(wsrc);
                while (w[0].y != (uint32_t)step) w[0] = ld_volatile_u2(wsrc);
            }
#pragma unroll
            for (int n = 1; n < 8; ++n)
                if (c_lo + n < c_hi) w[n] = ld_volatile_u2(wsrc + n * PS);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                if (c_lo + n < c_hi) {
                    while (w[n].y != (uint32_t)step) w[n] = ld_volatile_u2(wsrc + n * PS);
                    sx += __uint_as_float(w[n].x << 16);
                    sy += __uint_as_float(w[n].x & 0xffff0000u);
                }
            }
            s_part[(half * (NB / 2) + jp) * 32 + lane] = make_float2(sx, sy);
            __syncthreads();
            const float2 p0 = s_part[(0 * (NB / 2) + (j >> 1)) * 32 + unit], p1 = s_part[(1 * (NB / 2) + (j >> 1)) * 32 + unit];
            dh_rec = (j & 1) ? (p0.y + p1.y) : (p0.x + p1.x);
        }
        PROF_MARK(0);                                   // partial gather + reduce
        // ---- phase A: gate gradients ----
        const float gi = __uint_as_float(rg.x << 16), gf = __uint_as_float(rg.x & 0xffff0000u);
        const float gg = __uint_as_float(rg.y << 16), go = __uint_as_float(rg.y & 0xffff0000u);
        const float dh = __uint_as_float((uint32_t)rdy << 16) + dh_rec;
        const float tc_ = tanh_f(rc);
        const float dc = dc_carry + dh * go * (1.f - tc_ * tc_);
        const float da_i = dc * gg * gi * (1.f - gi);
        const float da_f = dc * rcp * gf * (1.f - gf);
        const float da_g = dc * gi * (1.f - gg * gg);
        const float da_o = dh * tc_ * go * (1.f - go);
        dc_carry = dc * gf;
        const float da[4] = {da_i, da_f, da_g, da_o};
        bf16 dab[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            dab[g] = __float2bfloat16_rn(da[g]);
            if (row_ok) bsum[g] += __bfloat162float(dab[g]);
            *reinterpret_cast<bf16 *>(sDA + tc::kmajor_off(j, g * 32 + unit, 128)) = dab[g];
        }
        auto store_and_prefetch = [&]() {                // global side effects of phase A, issued after the MMAs are in flight
            if (row_ok) {
                const uint32_t lo = (uint32_t)__bfloat16_as_ushort(dab[0]) | ((uint32_t)__bfloat16_as_ushort(dab[1]) << 16);
                const uint32_t hi = (uint32_t)__bfloat16_as_ushort(dab[2]) | ((uint32_t)__bfloat16_as_ushort(dab[3]) << 16);
                G2[g_off] = make_uint2(lo, hi);
            }
            g_off += g_step;
            y_off += y_step;
            if (step + 1 < T) {                          // raw prefetch for the next step
                rg = G2[g_off];
                rdy = dY16[y_off];
                rc = p.C[y_off];
                rcp = (step + 2 < T) ? p.C[y_off + y_step] : 0.f;
            }
        };
        if (step + 1 == T) { store_and_prefetch(); break; }     // dh_rec of the last step is never used
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
        PROF_MARK(1);                                   // gate gradients
        // ---- phase B: partial[jh, j] over my 128 gate rows ----
        if (warp < issuers) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const uint64_t b_desc0 = tc::smem_desc(tc::smem_u32(sDA), 128, 2048);
                for (int m = warp; m < tiles; m += issuers) {
#pragma unroll
                    for (int k16 = 0; k16 < 8; ++k16)
                        tc::mma_bf16_ts(tmem + dcol + m * NB, tmem + m * 64 + k16 * 8, b_desc0 + (uint64_t)(k16 * 16), idesc, k16 > 0);
                }
                tc::mma_commit(&s_bar);
            }
            __syncwarp();
        }
        store_and_prefetch();
        tc::mbar_wait(&s_bar, step & 1);
        tc::fence_after_sync();
        PROF_MARK(2);                                   // MMA
        // ---- phase C (producer side): ship partials to the owners of units jh ----
        uint2 *dst = p.ll + ((size_t)((step + 1) & 1) * groups + group) * ll_words;
        constexpr int kMaxTiles = 4;                     // H <= 512
        uint32_t v[kMaxTiles][CPW];
#pragma unroll
        for (int m = 0; m < kMaxTiles; ++m)              // all TMEM loads in flight, ONE wait
            if (m < tiles) tc::tmem_ld<CPW>(tmem + lane_base + dcol + m * NB + part * CPW, v[m]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int m = 0; m < kMaxTiles; ++m) {
            const int owner = 4 * m + q;                 // CTA that owns unit jh = 128 m + 32 q + lane
            if (m < tiles && owner < G) {
#pragma unroll
                for (int e = 0; e < CPW / 2; ++e) {
                    const __nv_bfloat162 pk = __floats2bfloat162_rn(__uint_as_float(v[m][2 * e]), __uint_as_float(v[m][2 * e + 1]));
                    const int jp = part * (CPW / 2) + e;
                    st_volatile_u2(dst + (((size_t)owner * G + u) * (NB / 2) + jp) * 32 + lane,
                                   make_uint2(*reinterpret_cast<const uint32_t *>(&pk), (uint32_t)(step + 1)));
                }
            }
        }
        tc::fence_before_sync();
        PROF_MARK(3);                                   // partial scatter
    }
    tc::fence_before_sync();
    __syncthreads();
    if (p.db_part) {
        // bias gradient of this CTA's 128 gate rows over its batch slice: fixed-order sum over the 16 rows (warps)
        float *s_b = reinterpret_cast<float *>(smem);                    // [16 rows][4 gates][32 units], reuses sDA + s_part
#pragma unroll
        for (int g = 0; g < 4; ++g) s_b[(warp * 4 + g) * 32 + lane] = bsum[g];
        __syncthreads();
        if (tid < 128) {
            const int g = tid >> 5, un = tid & 31;
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < NB; ++r) acc += s_b[(r * 4 + g) * 32 + un];
            p.db_part[((size_t)slice * 2 + d) * 4 * H + (size_t)g * H + u * kUnits + un] = acc;
        }
    }
    if (warp == 0) tc::tmem_dealloc(tmem, tmem_cols);
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

namespace {
int g_lstm_issuers = 2;      // tuning knob (mlvae_debug_set_option key 1)
int g_lstm_min_nb = 16;      // tuning knob (key 2): smallest batch slice per CTA
struct LstmPlan {
    int NB, slices, G;
    size_t smem, ll_bytes;
};
int lstm_plan(int B, int H, LstmPlan &pl) {
    MLVAE_REQUIRE(H % 32 == 0 && H >= 32 && H <= 512, MLVAE_ERR_UNSUPPORTED,
                  "lstm: hidden size must be a multiple of 32 in [32, 512] (W_hh slice resident in tensor memory), got %d", H);
    pl.G = H / kUnits;
    const int sms = sm_count();
    pl.NB = 16;                  // batch rows per CTA (MMA N); every CTA must be co-resident (1 CTA / SM)
    pl.slices = (B + pl.NB - 1) / pl.NB;
    MLVAE_REQUIRE((int64_t)pl.G * pl.slices * 2 <= sms, MLVAE_ERR_UNSUPPORTED,
                  "lstm: batch %d x hidden %d needs %d co-resident CTAs (> %d SMs)", B, H, pl.G * pl.slices * 2, sms);
    pl.smem = (size_t)pl.NB * H * 2;
    pl.ll_bytes = (size_t)2 * 2 * pl.slices * pl.G * pl.NB * 4 * sizeof(uint4);
    return MLVAE_OK;
}
}  // namespace

extern "C" {

// Debug: d_prof = 8 zeroed int64 cycle counters filled by CTA (0,0,0) of the next LSTM launches; NULL disables.
int mlvae_debug_set_profile_buffer(void *d_prof) {
    long long *ptr = (long long *)d_prof;
    MLVAE_CHECK_CUDA(cudaMemcpyToSymbol(g_prof, &ptr, sizeof(ptr)));
    return MLVAE_OK;
}

// Debug / tuning: key 1 = number of MMA issuer warps of the LSTM kernels (1, 2 or 4).
int mlvae_debug_set_option(int key, int value) {
    if (key == 1) { g_lstm_issuers = value; return MLVAE_OK; }
    if (key == 2 && (value == 16 || value == 32 || value == 64)) { g_lstm_min_nb = value; return MLVAE_OK; }
    return fail(MLVAE_ERR_INVALID_ARG, "unknown debug option %d", key);
}

// Scratch: the {data, tag} exchange words, zeroed by every call.
size_t mlvae_lstm_scratch_bytes(int B, int H) {
    LstmPlan pl;
    if (lstm_plan(B, H, pl) != MLVAE_OK) return 0;
    const size_t bwd = (size_t)2 * 2 * ((B + 15) / 16) * pl.G * pl.G * 8 * 32 * sizeof(uint2);
    return (pl.ll_bytes > bwd ? pl.ll_bytes : bwd) + 256;
}

int mlvae_lstm_fwd(void *d_p, const void *d_whh, void *d_y, float *d_c, int B, int T, int H, int save_gates,
                   void *d_scratch, void *stream) {
    MLVAE_REQUIRE(d_p && d_whh && d_y && d_scratch, MLVAE_ERR_INVALID_ARG, "lstm_fwd: missing buffers");
    MLVAE_REQUIRE(B > 0 && T > 0 && H > 0 && T < (1 << 30), MLVAE_ERR_INVALID_ARG, "lstm_fwd: bad sizes");
    LstmPlan pl;
    if (int rc = lstm_plan(B, H, pl)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    MLVAE_CHECK_CUDA(cudaMemsetAsync(d_scratch, 0, pl.ll_bytes, st));
    LstmFwdParams prm{(bf16 *)d_p, (const bf16 *)d_whh, (bf16 *)d_y, d_c, (uint4 *)d_scratch, B, T, H, save_gates};
    void *args[] = {&prm};
    dim3 grid(pl.G, pl.slices, 2), block(kLstmThreads);
    const void *fn = nullptr;
    const int nch = (g_lstm_issuers == 1 || g_lstm_issuers == 2 || g_lstm_issuers == 4) ? g_lstm_issuers : 2;
    fn = nch == 1 ? (const void *)lstm_fwd_kernel<16, 1>
                  : nch == 2 ? (const void *)lstm_fwd_kernel<16, 2> : (const void *)lstm_fwd_kernel<16, 4>;
    MLVAE_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    MLVAE_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, pl.smem, st));
    return MLVAE_OK;
}


int mlvae_lstm_bwd(void *d_gates, const float *d_c, const void *d_dy, const void *d_whh, float *d_bias_grad_part, int B, int T,
                   int H, void *d_scratch, void *stream) {
    MLVAE_REQUIRE(d_gates && d_c && d_dy && d_whh && d_scratch, MLVAE_ERR_INVALID_ARG, "lstm_bwd: missing buffers");
    MLVAE_REQUIRE(B > 0 && T > 0 && H > 0 && T < (1 << 30), MLVAE_ERR_INVALID_ARG, "lstm_bwd: bad sizes");
    LstmPlan pl;
    if (int rc = lstm_plan(B, H, pl)) return rc;
    const int slices = (B + 15) / 16;
    MLVAE_REQUIRE((int64_t)pl.G * slices * 2 <= sm_count(), MLVAE_ERR_UNSUPPORTED,
                  "lstm_bwd: batch %d x hidden %d needs %d co-resident CTAs", B, H, pl.G * slices * 2);
    const size_t ll_bytes = (size_t)2 * 2 * slices * pl.G * pl.G * 8 * 32 * sizeof(uint2);
    cudaStream_t st = (cudaStream_t)stream;
    MLVAE_CHECK_CUDA(cudaMemsetAsync(d_scratch, 0, ll_bytes, st));
    LstmBwdParams prm{(bf16 *)d_gates, d_c, (const bf16 *)d_dy, (const bf16 *)d_whh, (uint2 *)d_scratch, d_bias_grad_part, B, T, H};
    void *args[] = {&prm};
    dim3 grid(pl.G, slices, 2), block(kLstmThreads);
    const void *fn = (g_lstm_issuers == 1) ? (const void *)lstm_bwd_kernel<16, 1>
                     : (g_lstm_issuers == 4) ? (const void *)lstm_bwd_kernel<16, 4> : (const void *)lstm_bwd_kernel<16, 2>;
    const size_t smem = 16 * 4 * 32 * sizeof(float);       // >= sDA (4 KB) + s_part (4 KB); reused for the bias-gradient reduction
    MLVAE_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, smem, st));
    return MLVAE_OK;
}

}  // extern "C"
