// Persistent bidirectional LSTM recurrence for sm_100a (the decoder's
// 2 x biLSTM(512): modules/decoder.py:14-15,22).  cuDNN runs one GEMM + one cell kernel per
// timestep (2 x T launches per direction-layer); here ONE cooperative launch walks all T steps.
//
// Work split.  A CTA = (direction d, 16-row batch slice, 32-unit slice u) owns 128 gate rows of W_hh.
// That 128 x H bf16 slice is loaded ONCE into TENSOR MEMORY (H/2 32-bit columns) and is the A operand
// of every tcgen05.mma of the launch, so no weight byte moves during the sequence.  The input projection
// P = x W_ih^T + b_ih + b_hh for all timesteps is one big GEMM done before the launch.
//
// A step is a dependent chain  exchange h_{t-1} -> MMA -> gates -> publish h_t  whose cost is latency, not
// throughput.  What the round-2 measurements say about it (tests/lstm_probe.py, cycles at 1.97 GHz; the variants
// are listed in DESIGN.md section 4b):
//   * an L2 round trip under the polling load is ~600 cycles, so every EXTRA dependent load (round 1 spun on one
//     16-byte word, then fetched a second one) costs that much: each thread now polls exactly ONE 32-byte sector
//     (LDG.256) per step and validates all 16 element tags;
//   * the 32 K-steps of a step cost 0.8 k cycles when one thread issues them.  Independent accumulators should pipeline:
//     NOT so in practice -- one issuing thread manages one tcgen05.mma (M 128, N 16, A in TMEM) per ~26 cycles whatever
//     the accumulator (4 round-robin tiles, interleaved M-tiles: same 0.8 k cycles for 32 K-steps); two issuer WARPS do
//     run concurrently, so every chain has two, each with its own wave / accumulator tile (forward) or M-tiles (backward);
//   * MMA issue from inside `if (elect_one())` keeps the descriptors in per-thread registers (R2UR per operand,
//     ~17 cycles per MMA): the issuers are dedicated WARPS with uniform control flow that elect a lane only around
//     the tcgen05 instructions -> back-to-back UTCHMMA from uniform registers;
//   * the 16 batch rows of a CTA are TWO INDEPENDENT CHAINS of 8 rows, each run by its own 8 gate warps + 1 issuer
//     warp with NO synchronisation between the chains (no __syncthreads in the time loop; mbarrier hand-offs only):
//     a chain pulls 8 KB per step instead of 16 KB (the exchange time follows the bytes pulled), its gate phase has
//     2 warps per scheduler instead of 4 (that phase is issue bound), and while one chain waits on L2 the other
//     computes.  MMA N stays 16 (the minimum for M = 128): B-tile rows 8..15 are zeros, their columns never read.
//
// Exchange of h_t inside a group (the G = H/32 CTAs of one direction / chain) is a flag-in-data protocol
// over L2 (no fence / atomic / separate flag round trip): |h| <= 1 leaves bit 14 (the exponent MSB) of every
// bf16 free, and that bit of every element holds the step tag ((step + 1) >> 1) & 1 -- it alternates between
// successive uses of a slot and differs from the zeroed initial state.  Every element is validated on its own,
// so the protocol does not depend on store atomicity.  A NaN h (the only value with bit 14 set) travels as 0; the
// output Y keeps the NaN, so the loss is non-finite exactly when the reference's is.  The backward pass ships its
// partial dh sums the same way: scaled by 2^-64 (folded into the resident transposed weights, an exact power-of-two
// scaling) so that bit 14 is free as well (the sums are rescaled by 2^64 on arrival; gradients below 2^-62 flush to
// zero), 8 rows of two (producer, unit) pairs per 32-byte sector.  Deterministic: fixed summation orders, no data atomics.
// All CTAs wait on each other, so the launch is cooperative (co-residency guaranteed or refused).
#include <cooperative_groups.h>

#include "common.cuh"
#include "tc05.cuh"

namespace mlvae {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kUnits = 32;                   // hidden units per CTA
constexpr int kRows = 4 * kUnits;            // gate rows per CTA == MMA M
constexpr int kChains = 2;                   // independent 8-row recurrences per CTA
constexpr int kChainRows = 8;
constexpr int kGateWarps = 8;                // per chain
constexpr int kMmaN = 16;                    // MMA N (rows 8..15 of the B tiles are zero)
constexpr int kIssuers = 4;                  // MMA-issuer warps per chain, each with its own accumulator tile(s)
constexpr int kAcc = kIssuers;               // forward accumulator tiles per chain (one per issuer, added in the epilogue)
constexpr int kLstmThreads = (kChains * kGateWarps + kChains * kIssuers) * 32;      // 16 gate warps + 8 MMA-issuer warps = 768

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

struct u32x8 { uint4 lo, hi; };
// one 32-byte L2 sector per load (sm_100: LDG.256)
__device__ __forceinline__ u32x8 ld_volatile_u8(const uint4 *p) {
    u32x8 v;
    asm volatile("ld.relaxed.gpu.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v.lo.x), "=r"(v.lo.y), "=r"(v.lo.z), "=r"(v.lo.w), "=r"(v.hi.x), "=r"(v.hi.y), "=r"(v.hi.z), "=r"(v.hi.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u4(uint4 *p, uint4 v) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_volatile_u16(unsigned short *p, unsigned short v) {
    asm volatile("st.relaxed.gpu.global.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory");
}

// tag in bit 14 of every bf16 of the exchange words
constexpr uint32_t kTagBits = 0x40004000u;
__device__ __forceinline__ uint32_t step_tag(int step) { return (((step + 1) >> 1) & 1) ? kTagBits : 0u; }
__device__ __forceinline__ bool tag_ok(const u32x8 &w, uint32_t tag) {      // all 16 elements carry `tag` (0 or kTagBits)
    const uint32_t x = ((w.lo.x ^ tag) | (w.lo.y ^ tag) | (w.lo.z ^ tag) | (w.lo.w ^ tag)) |
                       ((w.hi.x ^ tag) | (w.hi.y ^ tag) | (w.hi.z ^ tag) | (w.hi.w ^ tag));
    return (x & kTagBits) == 0u;
}
__device__ __forceinline__ uint4 untag(const uint4 &w) { return make_uint4(w.x & ~kTagBits, w.y & ~kTagBits, w.z & ~kTagBits, w.w & ~kTagBits); }

__device__ long long *g_prof = nullptr;     // optional per-phase cycle counters (debug / profiles/)
__device__ int *g_trace = nullptr;          // optional per-STEP phase cycles of gate warp 0 / chain 0 / CTA (0,0,0): int[4 * T] (instrumented build)
// phase time stamps of ONE thread, accumulated in registers and written once at the end (a global read-modify-write per
// mark would put an L2 round trip into every phase)
// (the kernels are instantiated with and without the instrumentation: kProf = false compiles all of it away)
#define PROF_MARK(k)                                                     \
    do {                                                                 \
        if constexpr (kProf) { if (prof) { const long long now = clock64(); pacc[k] += now - tprev;                 \
            if (g_trace && prof == g_prof) g_trace[4 * step + (k)] = (int)(now - tprev);                             \
            tprev = now; } }                                                                                         \
    } while (0)
#define PROF_FLUSH()                                                     \
    do {                                                                 \
        if constexpr (kProf) { if (prof) { for (int k_ = 0; k_ < 4; ++k_) { prof[k_] += pacc[k_]; if (prof2) prof2[k_] += pacc[k_]; } } } \
    } while (0)
// profile buffer layout (int64, 128 entries): [0..3] gate warp 0 of chain 0 of CTA (0,0,0), [4..11] its issuer warps, [12 + 4 x ..] gate
// warp 0 of chain 0 of CTA (x,0,0) for every x (is one CTA of the group the slow one?), [76 + 4 w ..] gate warp w of chain 0 of CTA (0,0,0)
#define PROF_SETUP()                                                                                                                     \
    const bool prof_on = kProf && g_prof && blockIdx.y == 0 && blockIdx.z == 0 && chain == 0 && lane == 0 && (warp == 0 || blockIdx.x == 0); \
    long long *prof = !prof_on ? nullptr : (warp == 0 ? (blockIdx.x == 0 ? g_prof : g_prof + 12 + 4 * blockIdx.x) : g_prof + 76 + 4 * warp); \
    long long *prof2 = (prof_on && warp == 0 && blockIdx.x == 0) ? g_prof + 12 : nullptr

struct LstmFwdParams {
    bf16 *P;                 // (B, T, nd, H, 4) gate pre-activations (unit-major, the 4 gates i,f,g,o adjacent) from the input projection;
                             // overwritten with the ACTIVATED gates (i, f, g, o) when save != 0
    const bf16 *Whh;         // (nd, 4H, H)
    bf16 *Y;                 // (B, T, nd H)
    float *C;                // (B, T, nd H) cell states (saved for backward) or nullptr
    uint4 *ll;               // [2 parity][groups][G producers][4 quarters][8 rows] zeroed words of 8 tagged bf16
    int B, T, H, save;
    int poll_delay;          // cycles between publishing h_t and the first poll for the group's h_t (see g_lstm_poll_delay)
    int nd;                  // directions in the tensors' layout == gridDim.z == the kernel's kNd: 2 (bidirectional) or 1 (forward only)
};

template <bool kProf, int kNd>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_fwd_kernel(LstmFwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_full[kChains][kIssuers], s_mma[kChains], s_free[kChains];
    __shared__ uint32_t s_tmem;
    const int H = p.H, T = p.T, B = p.B;
    const int G = H / kUnits;                     // CTAs (producers) per group
    const int u = blockIdx.x;                     // unit slice (gridDim.x == G)
    const int slice = blockIdx.y;                 // 16-row batch slice
    const int d = blockIdx.z;                     // direction
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // broadcast: the compiler then keeps everything derived from it uniform
    const bool issuer = warp >= kChains * kGateWarps;
    const int chain = issuer ? (warp - kChains * kGateWarps) / kIssuers : warp / kGateWarps;
    const int iw = issuer ? (warp - kChains * kGateWarps) % kIssuers : 0;      // issuer warp inside the chain
    const int gw = warp % kGateWarps;             // gate warp inside the chain
    const int q = warp & 3;                       // TMEM lane quarter this warp may touch (hardware: warp % 4)
    const int part = gw >> 2;                     // batch rows part*4 .. part*4+3 of the chain
    const int b0 = (slice * kChains + chain) * kChainRows;
    const bool active = b0 < B;                   // whole groups agree on this (same slice, chain)
    // Gate rows are interleaved so that the 4 gates of a unit sit in 4 adjacent TMEM lanes:
    //   CTA row r = unit_local * 4 + gate  ->  lane = r % 32 of quarter r / 32
    // After the MMA a quad of lanes holds {i, f, g, o} x 4 batch rows of one unit; a 4 x 4 shuffle transpose gives
    // every lane all four gates of ONE (unit, batch row): no shared-memory round trip between MMA and cell update.
    const int gate = lane & 3;
    const int unit_local = 8 * q + (lane >> 2);

    const size_t tile_bytes = (size_t)kMmaN * H * 2;
    unsigned char *sH = smem + (size_t)chain * tile_bytes;                 // this chain's 16 x H bf16 K-major B tile

    const uint32_t dcol = (uint32_t)((H / 2 + 31) & ~31);                  // accumulator columns start
    uint32_t tmem_cols = 32;
    while (tmem_cols < dcol + kChains * kAcc * kMmaN) tmem_cols <<= 1;
    if (warp == 0) tc::tmem_alloc(&s_tmem, tmem_cols);
    if (tid == 0) {
        const int nw = (G + 1) / 2;                                          // gate warps per chain that pull words (2 producers each)
        for (int c = 0; c < kChains; ++c) {
            constexpr int wpw = kGateWarps / kIssuers;                      // gate warps per wave (wave i = warps i*wpw .. = producers 2*i*wpw ..)
            for (int i = 0; i < kIssuers; ++i) {
                const int n = nw - i * wpw;
                tc::mbar_init(&s_full[c][i], n <= 0 ? 1 : (n < wpw ? n : wpw));
            }
            tc::mbar_init(&s_mma[c], (nw + wpw - 1) / wpw);                  // one commit per issuer warp that has a wave
            tc::mbar_init(&s_free[c], kGateWarps);
        }
        tc::fence_barrier_init();
    }
    for (size_t i = (size_t)tid * 16; i < kChains * tile_bytes; i += (size_t)kLstmThreads * 16)
        *reinterpret_cast<uint4 *>(smem + i) = make_uint4(0, 0, 0, 0);     // rows 8..15 stay zero for the whole launch
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0);
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;

    // ---- resident weights: W_hh[d][gate*H + u*32 + unit_local][:] -> this thread's TMEM lane, columns k/2 ----
    if (!issuer) {
        const int wpart = warp >> 2;                                       // 4 warps per lane quarter share the K range
        const bf16 *wrow = p.Whh + ((size_t)d * 4 * H + (size_t)gate * H + u * kUnits + unit_local) * H;
        for (int k16 = wpart; k16 < H / 16; k16 += kChains * kGateWarps / 4) {
            const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(wrow + k16 * 16));
            const uint4 hi = __ldg(reinterpret_cast<const uint4 *>(wrow + k16 * 16 + 8));
            const uint32_t v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            tc::tmem_st8(tmem + lane_base + k16 * 8, v);
        }
        tc::tmem_st_wait();
    }
    tc::fence_proxy_async();                                               // the zero fill above -> async proxy
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    const uint32_t d_tile = tmem + dcol + chain * (kAcc * kMmaN);          // accumulator tile a at + a * kMmaN
    const int groups = kNd * gridDim.y * kChains;
    const int group = (d * gridDim.y + slice) * kChains + chain;
    const size_t ll_words = (size_t)G * 32;                                 // 16-byte words per group and parity

    if (issuer) {
        // ================= MMA issuer warps of this chain.  A single thread issues one tcgen05.mma per ~26 cycles whatever the
        // accumulator (measured), several warps issue concurrently: issuer iw owns wave iw (the K range of the producers that
        // gate warps 2 iw, 2 iw + 1 pull) and accumulator tile iw; the gate warps add the tiles.  The whole warp walks the sequence (uniform
        // control flow keeps the descriptors in uniform registers), one elected lane issues. =================
        constexpr int kpw = 2 * 2 * (kGateWarps / kIssuers);               // K-steps per wave: 2 per producer, 2 producers per gate warp
        const int k_lo = iw * kpw, k_hi = (k_lo + kpw) < 2 * G ? (k_lo + kpw) : 2 * G;
        if (active && k_lo < k_hi) {
            const uint32_t idesc = tc::idesc_bf16_f32(kRows, kMmaN);
            const uint64_t b_desc0 = tc::smem_desc(tc::smem_u32(sH), 128, (uint32_t)(H >> 3) * 128);
            const uint32_t my_tile = d_tile + iw * kMmaN;
            long long *prof = !kProf ? nullptr : ((g_prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && chain == 0 && lane == 0) ? g_prof + 4 + 2 * iw : nullptr);
            long long tprev = clock64(), pacc[4] = {0, 0, 0, 0};
            for (int step = 1; step < T; ++step) {
                if (step > 1) tc::mbar_wait(&s_free[chain], step & 1);     // the gate warps have read step-1's accumulators
                tc::mbar_wait(&s_full[chain][iw], (step - 1) & 1);
                tc::fence_after_sync();
                PROF_MARK(0);                                               // issuer: wait for its wave
                if (tc::elect_one()) {
#pragma unroll 8
                    for (int k = k_lo; k < k_hi; ++k)
                        tc::mma_bf16_ts(my_tile, tmem + k * 8, b_desc0 + (uint64_t)(k * 16), idesc, k > k_lo);
                    tc::mma_commit(&s_mma[chain]);
                }
                __syncwarp();
                PROF_MARK(1);                                               // issuer: issue + commit
            }
            if constexpr (kProf) { if (prof) { prof[0] += pacc[0]; prof[1] += pacc[1]; } }
        }
    } else if (active) {
        // ================= gate warps =================
        const int t_first = d ? (T - 1) : 0;
        constexpr int nd = kNd;                                              // directions in the tensors' layout (== gridDim.z)
        const ptrdiff_t p_step = (ptrdiff_t)(d ? -1 : 1) * nd * H;          // P stride per time step in 8-byte units
        const ptrdiff_t y_step = (ptrdiff_t)(d ? -1 : 1) * nd * H;
        const int my_row = part * 4 + gate;                                  // batch row this lane owns after the transpose
        const bool my_ok = (b0 + my_row) < B;
        const size_t y_off = ((size_t)min(b0 + my_row, B - 1) * T + t_first) * ((size_t)nd * H) + d * H + u * kUnits + unit_local;
        bf16 *pY = p.Y + y_off;
        float *pC = p.C ? p.C + y_off : nullptr;
        // gate buffer layout: (B, T, 2, H, 4) -- the four gates of a unit are adjacent, so the lane that owns (unit, batch
        // row) after the transpose reads its pre-activations and writes its activated gates with ONE 8-byte access
        uint2 *pG = reinterpret_cast<uint2 *>(p.P) +
                    (((size_t)min(b0 + my_row, B - 1) * T + t_first) * nd + d) * (size_t)H + u * kUnits + unit_local;
        // this warp's eight units of batch row my_row are one exchange word: producer u, row my_row, quarter q
        // exchange layout per group and parity: 16-byte word (producer, quarter, row) at (producer * 4 + quarter) * 8 + row, element e of
        // the word = unit 8 * quarter + e.  Every lane publishes its own 2 bytes: a warp (8 units x rows part*4 .. +3 of its quarter)
        // covers 64 CONTIGUOUS bytes = two whole 32-byte sectors with one store instruction (a sector filled by partial stores of two
        // warps takes two L2 transactions to become valid); assembling 16-byte words with shuffles first cost 3 dependent SHFL rounds
        // on the critical path (0.732 -> 0.711 ms per 500-step launch).
        // consumer side: thread i = gw*32 + lane pulls ONE sector per step: producer i / 16, quarter (i / 4) % 4, rows 2*(i % 4), +1
        // (8 units each).  Warp gw therefore pulls the 1 KB of producers 2gw, 2gw + 1.
        const int ci = gw * 32 + lane;
        const bool c_has = ci < G * 16;                                      // lane pulls a word
        const bool w_has = gw * 2 < G;                                       // warp pulls anything (warp-uniform)
        const int c_pr = ci >> 4, c_q = (ci >> 2) & 3, c_row = 2 * (ci & 3);
        const uint32_t c_soff = tc::kmajor_off(c_row, 8 * (4 * (c_has ? c_pr : 0) + c_q), H);   // second chunk (row + 1): + 16 bytes
        const int wave = gw / (kGateWarps / kIssuers);
        // exchange addresses of both parities, hoisted out of the time loop
        uint4 *const base0 = p.ll + (size_t)group * ll_words, *const base1 = base0 + (size_t)groups * ll_words;
        const size_t src_o = (size_t)2 * (c_has ? ci : 0), dst_o = ((size_t)u * 4 + q) * kChainRows + my_row;
        const uint4 *const src0 = base0 + src_o, *const src1 = base1 + src_o;
        unsigned short *const dst0 = reinterpret_cast<unsigned short *>(base0 + dst_o) + (lane >> 2);     // element (lane >> 2) of the word
        unsigned short *const dst1 = reinterpret_cast<unsigned short *>(base1 + dst_o) + (lane >> 2);
        const int nacc = ((G + 1) / 2 + kGateWarps / kIssuers - 1) / (kGateWarps / kIssuers);   // accumulator tiles in use (one per issuer with a wave)
        float c_state = 0.f;
        // input-projection terms are prefetched one step ahead as RAW bits (converting at load time would stall the warp
        // on the DRAM latency inside the step)
        uint2 pre_raw = *pG;

        PROF_SETUP();
        long long tprev = clock64(), pacc[4] = {0, 0, 0, 0};

        long long t_pub = 0;
        for (int step = 0; step < T; ++step) {
            float a[4];
            if (step > 0) {
                // ---- gather h_{t-1}: every thread polls ONE 32-byte sector (two exchange words) until all 16 element tags match ----
                if (w_has) {
                    if (c_has) {
                        const uint4 *src = (step & 1) ? src1 : src0;
                        const uint32_t tag = step_tag(step);
                        if (p.poll_delay > 0) while (clock64() - t_pub < p.poll_delay) {}   // a poll issued earlier cannot succeed: spare the L2
                        u32x8 w = ld_volatile_u8(src);
                        while (!tag_ok(w, tag)) w = ld_volatile_u8(src);
                        PROF_MARK(0);                       // exchange wait
                        *reinterpret_cast<uint4 *>(sH + c_soff) = untag(w.lo);
                        *reinterpret_cast<uint4 *>(sH + c_soff + 16) = untag(w.hi);
                        tc::fence_proxy_async();
                    }
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&s_full[chain][wave]);
                }
                PROF_MARK(1);                               // B-tile writes + hand-off
                tc::mbar_wait(&s_mma[chain], (step - 1) & 1);
                tc::fence_after_sync();
                PROF_MARK(2);                               // MMA completion (other warps' words, MMAs, commit)
                uint32_t v[kAcc][4];
#pragma unroll
                for (int t4 = 0; t4 < kAcc; ++t4) tc::tmem_ld<4>(lane_base + d_tile + t4 * kMmaN + part * 4, v[t4]);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float s = __uint_as_float(v[0][i]);
#pragma unroll
                    for (int t4 = 1; t4 < kAcc; ++t4)
                        if (t4 < nacc) s += __uint_as_float(v[t4][i]);
                    a[i] = s;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = 0.f;                     // h_0 = 0
            }
            // ---- 4 x 4 transpose of the raw accumulators inside each quad: lane `gate` ends up with the {i, f, g, o}
            //      pre-activations of batch row part*4 + gate for its unit ----
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
            // ---- cell update + publish ----
            const float c = a[1] * c_state + a[0] * a[2];
            c_state = c;
            const float h = a[3] * tanh_f(c);
            const bf16 hb = __float2bfloat16_rn(h);
            {
                uint32_t mine = (uint32_t)__bfloat16_as_ushort(hb);
                if (mine & 0x4000u) mine = 0;                                    // NaN (|h| <= 1 otherwise): travels as 0, Y keeps it
                mine |= step_tag(step + 1) & 0xffffu;
                // every lane stores its own 2 bytes: the warp's 32 lanes (8 units x rows part*4 .. +3) cover 64 CONTIGUOUS bytes = two
                // whole sectors with ONE store instruction, no shuffles to assemble words on the critical path
                if (step + 1 < T) st_volatile_u16((step & 1) ? dst0 : dst1, (unsigned short)mine);
                if (p.poll_delay > 0) t_pub = clock64();
            }
            // ---- everything below is off the critical path: it overlaps the L2 flight time of the words just published ----
            if (step > 0) {                                  // return the accumulators (their values left TMEM before the transpose)
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&s_free[chain]);
            }
            if (step + 1 < T) pre_raw = *(pG + p_step);      // prefetch next step's input projection (a full step to land)
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
            PROF_MARK(3);                                   // accumulator read + activations + cell update + publish + stores
        }
        PROF_FLUSH();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, tmem_cols);
}

// ======================================================================================
// Backward recurrence.  Same CTAs / chains / ownership as the forward pass.  Per step (reverse order), per chain:
//   consume  every thread pulls ONE sector: the 8 rows' partial dh sums one producer computed for two of this CTA's
//            units.  The 16 producers of a unit pair sit in the 16 lanes of a half warp: a reduce-scatter butterfly
//            (4 rounds, 11 SHFL, fixed tree = deterministic) leaves every lane the sum of ONE (unit, batch row) -- no
//            shared-memory round trip, no barrier across the chain's warps
//   phase A  thread = (unit 4 gw + lane / 8, batch row lane % 8): dh = dY + dh_rec; gate gradients -> pre-activation gradients
//            da_{i,f,g,o}; written (bf16) over the saved gates in G (input of the weight / input GEMMs that follow)
//            and into the chain's K-major B-operand tile (16 x 128 gate rows, rows 8..15 zero); mbarrier hand-off
//   phase B  partial[jh, j] = sum_{r in my 128 gate rows} W_hh[r, jh] * da[j, r]   tcgen05.mma with the TRANSPOSED
//            weight slice resident in TMEM (ceil(H/128) M-tiles x 64 columns), the M-tiles issued INTERLEAVED
//            (k outer, tile inner) so that consecutive MMAs hit different accumulators and pipeline
//   scatter  8-byte words (4 batch rows of one unit, scaled / tagged bf16, straight from the TMEM registers) to the
//            CTAs that own the units
// ======================================================================================
struct LstmBwdParams {
    bf16 *G;                 // (B, T, nd, H, 4) in: activated gates from the forward pass; out: pre-activation grads
    const float *C;          // (B, T, 2H) cell states from the forward pass
    const bf16 *dY;          // (B, T, 2H) gradient of the layer output
    const bf16 *Whh;         // (2, 4H, H)
    uint4 *ll;               // [2 parity][groups][G owners][8 unit quads][G producers][4 units] zeroed 16-byte words = 8 rows of tagged bf16
    float *db_part;          // (slices, nd, 4H) per-16-row-slice sums over (rows, t) of dA, torch gate order; or nullptr
    int B, T, H;
    int poll_delay;
    int nd;                  // directions (see LstmFwdParams)
};

constexpr float kWireUnscale = 18446744073709551616.f;    // 2^64
// The partial dh sums travel as bf16 scaled by 2^-64 so that they stay below 2.0 (bit 14 of the bf16 free for the tag).  The scale
// is folded into the resident weights (exact: a power of two), so the accumulators come out of TMEM already scaled and packing a
// pair is one convert + one logic op on the critical path.  Weights below 2^-62 flush to zero.
__device__ __forceinline__ uint32_t bf16_scale_down64(uint32_t x) {      // bf16 bit pattern * 2^-64
    return ((x >> 7) & 0xffu) > 64u ? x - (64u << 7) : (x & 0x8000u);
}
__device__ __forceinline__ uint32_t wire_pack(float x0, float x1, uint32_t tag) {
    // round to bf16, force the tag bit.  A non-finite sum (exponent MSB set) becomes a huge finite value on arrival; the NaN / Inf is
    // already recorded in dA of this row / step, which is what makes the weight gradients non-finite.  Finite sums of 2^65 and
    // more (no training run survives those) are not representable on the wire.
    const __nv_bfloat162 pk = __floats2bfloat162_rn(x0, x1);
    return (*reinterpret_cast<const uint32_t *>(&pk) & ~kTagBits) | tag;
}

template <bool kProf, int kNd>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_bwd_kernel(LstmBwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_da[kChains][2], s_mma[kChains], s_free[kChains];
    __shared__ uint32_t s_tmem;
    const int H = p.H, T = p.T, B = p.B;
    const int G = H / kUnits;
    const int u = blockIdx.x, slice = blockIdx.y, d = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const bool issuer = warp >= kChains * kGateWarps;
    const int chain = issuer ? (warp - kChains * kGateWarps) / kIssuers : warp / kGateWarps;
    const int iw = issuer ? (warp - kChains * kGateWarps) % kIssuers : 0;
    const int gw = warp % kGateWarps;
    const int q = warp & 3, part = gw >> 2;
    const int b0 = (slice * kChains + chain) * kChainRows;
    const bool active = b0 < B;
    const int tiles = (H + 127) / 128;

    constexpr int kDaBytes = kMmaN * 128 * 2;                                        // 4 KB per chain
    unsigned char *sDA = smem + chain * kDaBytes;                                    // 16 x 128 bf16, K-major

    const uint32_t dcol = (uint32_t)((tiles * 64 + 31) & ~31);
    uint32_t tmem_cols = 32;
    while (tmem_cols < dcol + (uint32_t)kChains * 4 * kMmaN) tmem_cols <<= 1;
    if (warp == 0) tc::tmem_alloc(&s_tmem, tmem_cols);
    if (tid == 0) {
        for (int c = 0; c < kChains; ++c) {
            tc::mbar_init(&s_da[c][0], kGateWarps / 2);      // wave 0: gate warps 0..3 = units 0..15 = the even K-steps
            tc::mbar_init(&s_da[c][1], kGateWarps / 2);      // wave 1: gate warps 4..7 = units 16..31 = the odd K-steps
            tc::mbar_init(&s_mma[c], tiles < kIssuers ? tiles : kIssuers);       // one commit per issuer warp that owns M-tiles
            tc::mbar_init(&s_free[c], kGateWarps);
        }
        tc::fence_barrier_init();
    }
    for (int i = tid * 16; i < kChains * kDaBytes; i += kLstmThreads * 16) *reinterpret_cast<uint4 *>(smem + i) = make_uint4(0, 0, 0, 0);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem, 0);
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;

    // ---- transposed weight slice -> TMEM: tile m, lane = jh - 128 m, K index r = gate*32 + unit ----
    if (!issuer) {
        const unsigned short *W16 = reinterpret_cast<const unsigned short *>(p.Whh) + (size_t)d * 4 * H * H;
        const int wpart = warp >> 2;
        for (int m = wpart; m < tiles; m += kChains * kGateWarps / 4) {
            const int jh = 128 * m + 32 * q + lane;
            for (int k16 = 0; k16 < 8; ++k16) {
                uint32_t v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    uint32_t lo = 0, hi = 0;
                    if (jh < H) {
                        const int r0 = k16 * 16 + 2 * e, r1 = r0 + 1;
                        lo = bf16_scale_down64(W16[((size_t)(r0 >> 5) * H + u * kUnits + (r0 & 31)) * H + jh]);
                        hi = bf16_scale_down64(W16[((size_t)(r1 >> 5) * H + u * kUnits + (r1 & 31)) * H + jh]);
                    }
                    v[e] = lo | (hi << 16);
                }
                tc::tmem_st8(tmem + lane_base + m * 64 + k16 * 8, v);
            }
        }
        tc::tmem_st_wait();
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    const int groups = kNd * gridDim.y * kChains;
    const int group = (d * gridDim.y + slice) * kChains + chain;
    const size_t ll_words = (size_t)G * G * 32;                                      // 16-byte words per group and parity
    const uint32_t d_tile0 = tmem + dcol + chain * 4 * kMmaN;

    float bsum[4] = {0.f, 0.f, 0.f, 0.f};             // bias gradient: sum over t of this (row, unit)'s dA, per gate

    if (issuer) {
        // issuer iw owns the M-tiles iw, iw + 2 (two warps issue concurrently; one thread manages one tcgen05.mma per ~26 cycles)
        if (active && iw < tiles) {
            const uint32_t idesc = tc::idesc_bf16_f32(128, kMmaN);
            const uint64_t b_desc0 = tc::smem_desc(tc::smem_u32(sDA), 128, 2048);
            long long *prof = !kProf ? nullptr : ((g_prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && chain == 0 && lane == 0) ? g_prof + 4 + 2 * iw : nullptr);
            long long tprev = clock64(), pacc[4] = {0, 0, 0, 0};
            for (int step = 0; step + 1 < T; ++step) {                               // the last step ships no partials
                if (step > 0) tc::mbar_wait(&s_free[chain], (step - 1) & 1);         // previous partials have left TMEM
                // K index = gate * 32 + unit: the even K-steps hold units 0..15 (gate warps 0..3), the odd ones units 16..31
                tc::mbar_wait(&s_da[chain][0], step & 1);
                tc::fence_after_sync();
                PROF_MARK(0);                                                        // issuer: wait for dA
                if (tc::elect_one()) {
                    for (int m = iw; m < tiles; m += kIssuers) {
#pragma unroll
                        for (int k16 = 0; k16 < 8; k16 += 2)
                            tc::mma_bf16_ts(d_tile0 + m * kMmaN, tmem + m * 64 + k16 * 8, b_desc0 + (uint64_t)(k16 * 16), idesc, k16 > 0);
                    }
                }
                __syncwarp();
                tc::mbar_wait(&s_da[chain][1], step & 1);
                tc::fence_after_sync();
                if (tc::elect_one()) {
                    for (int m = iw; m < tiles; m += kIssuers) {
#pragma unroll
                        for (int k16 = 1; k16 < 8; k16 += 2)
                            tc::mma_bf16_ts(d_tile0 + m * kMmaN, tmem + m * 64 + k16 * 8, b_desc0 + (uint64_t)(k16 * 16), idesc, 1);
                    }
                    tc::mma_commit(&s_mma[chain]);
                }
                __syncwarp();
                PROF_MARK(1);                                                        // issuer: issue + commit
            }
            if constexpr (kProf) { if (prof) { prof[0] += pacc[0]; prof[1] += pacc[1]; } }
        }
    } else if (active) {
        // phase-A identity of this thread = where the reduce-scatter leaves its sum: unit 4 gw + lane / 8, batch row lane % 8
        const int j = lane & 7, unit = 4 * gw + (lane >> 3);
        const int b = min(b0 + j, B - 1);                 // rows past B are clamped for loads, never stored
        const bool row_ok = (b0 + j) < B;
        uint2 *G2 = reinterpret_cast<uint2 *>(p.G);
        const unsigned short *dY16 = reinterpret_cast<const unsigned short *>(p.dY);
        const int t_first = d ? 0 : (T - 1);
        constexpr int nd = kNd;                                              // directions in the tensors' layout (== gridDim.z)
        const ptrdiff_t g_step = (ptrdiff_t)(d ? 1 : -1) * nd * H;
        const ptrdiff_t y_step = (ptrdiff_t)(d ? 1 : -1) * nd * H;
        size_t g_off = (((size_t)b * T + t_first) * nd + d) * (size_t)H + u * kUnits + unit;       // 8-byte units: (B,T,nd,H,4)
        size_t y_off = ((size_t)b * T + t_first) * ((size_t)nd * H) + d * H + u * kUnits + unit;
        // consumer identity: lane pulls the sector of producer lane % 16, units 4 gw + 2 (lane / 16), + 1: the 16 producers of a unit
        // pair are the 16 lanes of a half warp
        const int c_pr = lane & 15;
        const bool c_has = c_pr < G;
        // producer identity: the warps with part == 0 ship the M-tiles 0, 1 and those with part == 1 the tiles 2, 3, each thread ALL
        // eight rows of unit `lane` of owner 4m+q: one 16-byte word per thread, 512 contiguous bytes (16 whole sectors) per warp store
        uint4 *const base0 = p.ll + (size_t)group * ll_words, *const base1 = base0 + (size_t)groups * ll_words;
#ifndef MLVAE_LSTM_BWD_LAYOUT_V1
        // exchange layout per group and parity: word [owner][unit quad = unit / 4][producer][unit % 4]: the 32 sectors a consumer warp
        // polls (unit quad gw, all producers) are ONE contiguous kilobyte = 8 whole 128-byte lines per poll instruction (with the
        // producer-major layout [owner][producer][unit] they were 16 half lines, twice the L2 request count under the polling load);
        // a producer warp store covers 8 x 64 contiguous bytes (whole sectors)
        const size_t src_o = (((size_t)u * 8 + gw) * G + (c_has ? c_pr : 0)) * 4 + 2 * (lane >> 4);
        const size_t dst_o = ((size_t)(lane >> 2) * G + u) * 4 + (lane & 3);       // + owner * G * 32
#else      // producer-major layout [owner][producer][unit] (A/B only: 16 half lines per poll; 4 % slower on the GPUs with less L2 request head-room)
        const size_t src_o = (size_t)u * G * 32 + (size_t)(c_has ? c_pr : 0) * 32 + 4 * gw + 2 * (lane >> 4), dst_o = (size_t)u * 32 + lane;
#endif
        const uint4 *const src0 = base0 + src_o, *const src1 = base1 + src_o;
        uint4 *const dst0 = base0 + dst_o, *const dst1 = base1 + dst_o;       // + owner * G * 32

        float dc_carry = 0.f;
        uint2 rg = G2[g_off];
        unsigned short rdy = dY16[y_off];
        float rc = p.C[y_off];
        float rcp = (T > 1) ? p.C[y_off + y_step] : 0.f;          // c_{t-1} in forward order == next time index visited here

        PROF_SETUP();
        long long tprev = clock64(), pacc[4] = {0, 0, 0, 0};

        long long t_pub = 0;
        for (int step = 0; step < T; ++step) {
            float dh_rec = 0.f;
            if (step > 0) {
                // ---- consume: the partial sums addressed to this CTA, one sector per thread ----
                uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
                if (c_has) {
                    const uint4 *src = (step & 1) ? src1 : src0;
                    const uint32_t tag = step_tag(step);
                    if (p.poll_delay > 0) while (clock64() - t_pub < p.poll_delay) {}
                    u32x8 w = ld_volatile_u8(src);
                    while (!tag_ok(w, tag)) w = ld_volatile_u8(src);
                    lo = untag(w.lo);
                    hi = untag(w.hi);
                }
                PROF_MARK(0);                               // exchange wait
                {
                    // lo = rows 0..7 of unit c_u0, hi = rows 0..7 of unit c_u0 + 1 (bf16 pairs) from producer lane % 16.
                    // Reduce-scatter over the 16 producers: round 1 (xor 8) trades whole packed words and settles the unit,
                    // rounds 2..4 (xor 4, 2, 1) halve the rows; both partners of a pair add the same two numbers.
                    const bool k1 = lane & 8, k2 = lane & 4, k3 = lane & 2, k4 = lane & 1;
                    const uint4 mine = k1 ? hi : lo, send = k1 ? lo : hi;
                    uint4 got;
                    got.x = __shfl_xor_sync(0xffffffffu, send.x, 8);
                    got.y = __shfl_xor_sync(0xffffffffu, send.y, 8);
                    got.z = __shfl_xor_sync(0xffffffffu, send.z, 8);
                    got.w = __shfl_xor_sync(0xffffffffu, send.w, 8);
                    const uint32_t a4[4] = {mine.x, mine.y, mine.z, mine.w}, b4[4] = {got.x, got.y, got.z, got.w};
                    float f[8];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        f[2 * r] = __uint_as_float(a4[r] << 16) + __uint_as_float(b4[r] << 16);
                        f[2 * r + 1] = __uint_as_float(a4[r] & 0xffff0000u) + __uint_as_float(b4[r] & 0xffff0000u);
                    }
                    float g4[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float keep = k2 ? f[4 + r] : f[r], give = k2 ? f[r] : f[4 + r];
                        g4[r] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
                    }
                    float g2[2];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const float keep = k3 ? g4[2 + r] : g4[r], give = k3 ? g4[r] : g4[2 + r];
                        g2[r] = keep + __shfl_xor_sync(0xffffffffu, give, 2);
                    }
                    const float keep = k4 ? g2[1] : g2[0], give = k4 ? g2[0] : g2[1];
                    dh_rec = (keep + __shfl_xor_sync(0xffffffffu, give, 1)) * kWireUnscale;      // unit 4 gw + lane / 8, row lane % 8
                }
                PROF_MARK(1);                               // reduce-scatter inside the warp
            }
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
            }
            const bool last = step + 1 == T;
            if (!last) {
#pragma unroll
                for (int g = 0; g < 4; ++g) *reinterpret_cast<bf16 *>(sDA + tc::kmajor_off(j, g * 32 + unit, 128)) = dab[g];
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&s_da[chain][gw >> 2]);
            }
            // global side effects of phase A, after the hand-off to the issuer
            if (row_ok) {
                const uint32_t lo = (uint32_t)__bfloat16_as_ushort(dab[0]) | ((uint32_t)__bfloat16_as_ushort(dab[1]) << 16);
                const uint32_t hi = (uint32_t)__bfloat16_as_ushort(dab[2]) | ((uint32_t)__bfloat16_as_ushort(dab[3]) << 16);
                G2[g_off] = make_uint2(lo, hi);
            }
            g_off += g_step;
            y_off += y_step;
            if (!last) {                                 // raw prefetch for the next step
                rg = G2[g_off];
                rdy = dY16[y_off];
                rc = p.C[y_off];
                rcp = (step + 2 < T) ? p.C[y_off + y_step] : 0.f;
            }
            PROF_MARK(2);                               // gate gradients + hand-off
            if (last) break;
            // ---- scatter ----
            uint4 *dst = (step & 1) ? dst0 : dst1;
            const uint32_t tg_out = step_tag(step + 1);
            tc::mbar_wait(&s_mma[chain], step & 1);
            tc::fence_after_sync();
            uint32_t v[2][8];
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (part * 2 + i < tiles) tc::tmem_ld<8>(lane_base + d_tile0 + (part * 2 + i) * kMmaN, v[i]);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int m = part * 2 + i;
                const int owner = 4 * m + q;                 // CTA that owns unit jh = 128 m + 32 q + lane
                if (m < tiles && owner < G)
                    st_volatile_u4(dst + (size_t)owner * G * 32,
                                   make_uint4(wire_pack(__uint_as_float(v[i][0]), __uint_as_float(v[i][1]), tg_out),
                                              wire_pack(__uint_as_float(v[i][2]), __uint_as_float(v[i][3]), tg_out),
                                              wire_pack(__uint_as_float(v[i][4]), __uint_as_float(v[i][5]), tg_out),
                                              wire_pack(__uint_as_float(v[i][6]), __uint_as_float(v[i][7]), tg_out)));
            }
            if (p.poll_delay > 0) t_pub = clock64();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s_free[chain]);
            PROF_MARK(3);                               // MMA completion + scatter
        }
        PROF_FLUSH();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (p.db_part) {
        // bias gradient of this CTA's 128 gate rows over its 16-row slice: fixed-order sum over the 16 (chain, row) warps
        float *s_b = reinterpret_cast<float *>(smem);                    // [16 rows][4 gates][32 units], reuses the operand tiles
        if (!issuer) {
            const int r = chain * kChainRows + (lane & 7), un = 4 * gw + (lane >> 3);      // this thread's (row of the slice, unit)
#pragma unroll
            for (int g = 0; g < 4; ++g) s_b[(r * 4 + g) * 32 + un] = active ? bsum[g] : 0.f;
        }
        __syncthreads();
        if (tid < 128) {
            const int g = tid >> 5, un = tid & 31;
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < kChains * kGateWarps; ++r) acc += s_b[(r * 4 + g) * 32 + un];
            p.db_part[((size_t)slice * kNd + d) * 4 * H + (size_t)g * H + u * kUnits + un] = acc;
        }
    }
    if (warp == 0) tc::tmem_dealloc(tmem, tmem_cols);
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

namespace {
// A poll issued before the peers' words can have landed only loads the L2 (and every failed poll costs a ~600-cycle round trip):
// the gate warps wait this many cycles after publishing before their first poll.  Sweeps on B200 (tests/probes/lstm_kernel_times.py,
// ms per 500-step launch): forward 0 / 200 / 400 / 600 / 800 / 1000 / 1300 cycles -> 0.838 / 0.837 / 0.812 / 0.785 / 0.801 / 0.842 / 0.938;
// backward (publishes at the very end of its step; in-warp reduce, no chain barrier) 0 / 300 / 500 / 700 / 900 -> 0.850 / 0.840 /
// 0.837 / 0.865 / 0.895 (a second box: 0.972 / 0.965 / 0.938 at 0 / 300 / 500).  GPUs of the pool fall into two classes for the backward
// kernel (profiles/r02_per_gpu_spread.txt): 0.83-0.84 ms at 400 and +2 % at 600 on most, 0.91-0.92 ms at 400 and -1.5 % at 600 on the
// others (less L2 request head-room: they also lose 4 % with the half-line polling layout MLVAE_LSTM_BWD_LAYOUT_V1); 500 suits both.
int g_lstm_poll_delay_fwd = 600, g_lstm_poll_delay_bwd = 500;
bool g_lstm_prof = false;        // a profile buffer is set: launch the instrumented instantiations
struct LstmPlan {
    int slices, G;
    size_t smem_fwd, smem_bwd, ll_fwd, ll_bwd;
};
int lstm_plan(int B, int H, int nd, LstmPlan &pl) {
    MLVAE_REQUIRE(H % 32 == 0 && H >= 32 && H <= 512, MLVAE_ERR_UNSUPPORTED,
                  "lstm: hidden size must be a multiple of 32 in [32, 512] (W_hh slice resident in tensor memory), got %d", H);
    MLVAE_REQUIRE(nd == 1 || nd == 2, MLVAE_ERR_INVALID_ARG, "lstm: 1 or 2 directions, got %d", nd);
    pl.G = H / kUnits;
    const int sms = sm_count();
    pl.slices = (B + kChains * kChainRows - 1) / (kChains * kChainRows);          // 16 batch rows (two 8-row chains) per CTA
    MLVAE_REQUIRE((int64_t)pl.G * pl.slices * nd <= sms, MLVAE_ERR_UNSUPPORTED,
                  "lstm: batch %d x hidden %d x %d direction(s) needs %d co-resident CTAs (> %d SMs)", B, H, nd, pl.G * pl.slices * nd, sms);
    const size_t groups = (size_t)nd * pl.slices * kChains;
    pl.smem_fwd = (size_t)kChains * kMmaN * H * 2;
    pl.smem_bwd = (size_t)kChains * kMmaN * 128 * 2;                                // dA tiles = 8 KB (also holds the bias reduction's 8 KB)
    pl.ll_fwd = 2 * groups * pl.G * 32 * sizeof(uint4);
    pl.ll_bwd = 2 * groups * (size_t)pl.G * pl.G * 32 * sizeof(uint4);
    return MLVAE_OK;
}
}  // namespace

extern "C" {

// Debug: d_prof = 128 zeroed int64 cycle counters (layout at PROF_SETUP) filled by the next LSTM launches; NULL disables.
int mlvae_debug_set_profile_buffer(void *d_prof) {
    long long *ptr = (long long *)d_prof;
    MLVAE_CHECK_CUDA(cudaMemcpyToSymbol(g_prof, &ptr, sizeof(ptr)));
    g_lstm_prof = ptr != nullptr;
    return MLVAE_OK;
}

// Debug: d_trace = 4 * T ints receiving the four phase durations of EVERY step of gate warp 0, chain 0, CTA (0,0,0) (needs a profile
// buffer as well: the instrumented instantiation); NULL disables.
int mlvae_debug_set_trace_buffer(void *d_trace) {
    int *ptr = (int *)d_trace;
    MLVAE_CHECK_CUDA(cudaMemcpyToSymbol(g_trace, &ptr, sizeof(ptr)));
    return MLVAE_OK;
}

// Debug / tuning knobs: key 3 = cycles a gate warp waits after publishing before its first poll (forward), key 4 = same for
// the backward kernel.
int mlvae_debug_set_option(int key, int value) {
    if (key == 3 && value >= 0) { g_lstm_poll_delay_fwd = value; return MLVAE_OK; }
    if (key == 4 && value >= 0) { g_lstm_poll_delay_bwd = value; return MLVAE_OK; }
    return fail(MLVAE_ERR_INVALID_ARG, "unknown debug option %d", key);
}

// Scratch: the tagged exchange words, zeroed by every call.
size_t mlvae_lstm_scratch_bytes_dirs(int B, int H, int ndir) {
    LstmPlan pl;
    if (B <= 0 || lstm_plan(B, H, ndir, pl) != MLVAE_OK) return 0;
    return (pl.ll_fwd > pl.ll_bwd ? pl.ll_fwd : pl.ll_bwd) + 256;
}
size_t mlvae_lstm_scratch_bytes(int B, int H) { return mlvae_lstm_scratch_bytes_dirs(B, H, 2); }

int mlvae_lstm_fwd_dirs(void *d_p, const void *d_whh, void *d_y, float *d_c, int B, int T, int H, int ndir, int save_gates,
                        void *d_scratch, void *stream) {
    MLVAE_REQUIRE(d_p && d_whh && d_y && d_scratch, MLVAE_ERR_INVALID_ARG, "lstm_fwd: missing buffers");
    MLVAE_REQUIRE(B > 0 && T > 0 && H > 0 && T < (1 << 30), MLVAE_ERR_INVALID_ARG, "lstm_fwd: bad sizes");
    MLVAE_REQUIRE(((uintptr_t)d_scratch & 31) == 0, MLVAE_ERR_INVALID_ARG, "lstm_fwd: scratch must be 32-byte aligned");
    LstmPlan pl;
    if (int rc = lstm_plan(B, H, ndir, pl)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    MLVAE_CHECK_CUDA(cudaMemsetAsync(d_scratch, 0, pl.ll_fwd, st));
    LstmFwdParams prm{(bf16 *)d_p, (const bf16 *)d_whh, (bf16 *)d_y, d_c, (uint4 *)d_scratch, B, T, H, save_gates, g_lstm_poll_delay_fwd, ndir};
    void *args[] = {&prm};
    dim3 grid(pl.G, pl.slices, ndir), block(kLstmThreads);
    const void *fn = ndir == 2 ? (g_lstm_prof ? (const void *)lstm_fwd_kernel<true, 2> : (const void *)lstm_fwd_kernel<false, 2>)
                               : (g_lstm_prof ? (const void *)lstm_fwd_kernel<true, 1> : (const void *)lstm_fwd_kernel<false, 1>);
    MLVAE_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_fwd));
    MLVAE_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, pl.smem_fwd, st));
    return MLVAE_OK;
}

int mlvae_lstm_fwd(void *d_p, const void *d_whh, void *d_y, float *d_c, int B, int T, int H, int save_gates,
                   void *d_scratch, void *stream) {
    return mlvae_lstm_fwd_dirs(d_p, d_whh, d_y, d_c, B, T, H, 2, save_gates, d_scratch, stream);
}

int mlvae_lstm_bwd_dirs(void *d_gates, const float *d_c, const void *d_dy, const void *d_whh, float *d_bias_grad_part, int B, int T,
                        int H, int ndir, void *d_scratch, void *stream) {
    MLVAE_REQUIRE(d_gates && d_c && d_dy && d_whh && d_scratch, MLVAE_ERR_INVALID_ARG, "lstm_bwd: missing buffers");
    MLVAE_REQUIRE(B > 0 && T > 0 && H > 0 && T < (1 << 30), MLVAE_ERR_INVALID_ARG, "lstm_bwd: bad sizes");
    MLVAE_REQUIRE(((uintptr_t)d_scratch & 31) == 0, MLVAE_ERR_INVALID_ARG, "lstm_bwd: scratch must be 32-byte aligned");
    LstmPlan pl;
    if (int rc = lstm_plan(B, H, ndir, pl)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    MLVAE_CHECK_CUDA(cudaMemsetAsync(d_scratch, 0, pl.ll_bwd, st));
    LstmBwdParams prm{(bf16 *)d_gates, d_c, (const bf16 *)d_dy, (const bf16 *)d_whh, (uint4 *)d_scratch, d_bias_grad_part, B, T, H, g_lstm_poll_delay_bwd, ndir};
    void *args[] = {&prm};
    dim3 grid(pl.G, pl.slices, ndir), block(kLstmThreads);
    const void *fn = ndir == 2 ? (g_lstm_prof ? (const void *)lstm_bwd_kernel<true, 2> : (const void *)lstm_bwd_kernel<false, 2>)
                               : (g_lstm_prof ? (const void *)lstm_bwd_kernel<true, 1> : (const void *)lstm_bwd_kernel<false, 1>);
    MLVAE_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, pl.smem_bwd, st));
    return MLVAE_OK;
}

int mlvae_lstm_bwd(void *d_gates, const float *d_c, const void *d_dy, const void *d_whh, float *d_bias_grad_part, int B, int T,
                   int H, void *d_scratch, void *stream) {
    return mlvae_lstm_bwd_dirs(d_gates, d_c, d_dy, d_whh, d_bias_grad_part, B, T, H, 2, d_scratch, stream);
}

}  // extern "C"
