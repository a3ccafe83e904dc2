// Persistent bidirectional LSTM recurrence for sm_100a (the decoder's
// 2 x biLSTM(512): modules/decoder.py:14-15,22).  cuDNN runs one GEMM + one cell kernel per
// timestep (2 x T launches per direction-layer); here ONE cooperative launch walks all T steps.
//
// Work split.  A "group" = (direction d, batch slice s of NB rows).  Its G = H/32 CTAs each own 32
// hidden units = 128 gate rows (row = gate*32 + unit) of W_hh, resident in shared memory for the
// whole sequence (128 x H bf16, K-major).  Per step every CTA computes
//     D[128 gate rows x NB batch] = W_slice (smem) x h_{t-1}^T (smem)        tcgen05.mma, fp32 in TMEM
// adds the precomputed input projection P[b, t] (x W_ih^T + b_ih + b_hh, one big GEMM done before),
// applies the gate non-linearities, updates c (registers) and writes its 32-unit slice of h_t straight
// into the output tensor Y[b, t, d*H + units].  Y doubles as the exchange buffer: the group's CTAs
// publish "step done" through one global counter (release: __syncthreads + __threadfence + atomicAdd,
// acquire: ld.acquire.gpu poll) and then gather h_t (NB x H bf16 = 16 KB at NB = 16) from L2 with
// cp.async into the K-major B-operand tile.  No data-path atomics, deterministic.
// All CTAs must be co-resident (they wait on each other): the launch is cooperative.
#include <cooperative_groups.h>

#include "common.cuh"
#include "tc05.cuh"

namespace mlvae {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kUnits = 32;            // hidden units per CTA
constexpr int kRows = 4 * kUnits;     // gate rows per CTA == MMA M
constexpr int kLstmThreads = 128;

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return 2.f / (1.f + __expf(-2.f * x)) - 1.f; }

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ long long *g_prof = nullptr;     // optional per-phase cycle counters (debug / profiles/)
#define PROF_MARK(k)                                                     \
    do {                                                                 \
        if (prof && tid == 0) { const long long now = clock64(); prof[k] += now - tprev; tprev = now; } \
    } while (0)

struct LstmFwdParams {
    bf16 *P;                 // (B, T, 2, 4H) gate pre-activations from the input projection (+ biases);
                             // overwritten with the ACTIVATED gates (i, f, g, o) when save != 0
    const bf16 *Whh;         // (2, 4H, H)
    bf16 *Y;                 // (B, T, 2H)
    float *C;                // (B, T, 2H) cell states (saved for backward) or nullptr
    unsigned int *flags;     // (2 * n_slices) zeroed counters
    int B, T, H, save;
};

template <int NB>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_fwd_kernel(LstmFwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int H = p.H, T = p.T, B = p.B;
    const int G = H / kUnits;
    const int u = blockIdx.x;                     // unit slice (gridDim.x == G)
    const int slice = blockIdx.y;                 // batch slice
    const int d = blockIdx.z;                     // direction
    const int b0 = slice * NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    unsigned char *sW = smem;                                  // 128 x H bf16, K-major
    unsigned char *sH = smem + (size_t)kRows * H * 2;          // NB x H bf16, K-major
    float *s_act = reinterpret_cast<float *>(sH + (size_t)NB * H * 2);   // [4][NB][32]

    if (warp == 0) tc::tmem_alloc(&s_tmem, NB < 32 ? 32 : NB);
    if (tid == 0) {
        tc::mbar_init(&s_bar, 1);
        tc::fence_barrier_init();
    }
    // ---- resident weight slice: row r = gate*32 + unit  <-  W_hh[d][gate*H + u*32 + unit][:] ----
    {
        const int chunks = H >> 3;
        const bf16 *Wd = p.Whh + (size_t)d * 4 * H * H;
        for (int i = tid; i < kRows * chunks; i += kLstmThreads) {
            const int r = (i & 7) | ((i / (8 * chunks)) << 3);
            const int c = (i >> 3) % chunks;
            const int grow = (r >> 5) * H + u * kUnits + (r & 31);
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(Wd + (size_t)grow * H + c * 8));
            *reinterpret_cast<uint4 *>(sW + tc::kmajor_off(r, c * 8, H)) = v;
        }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = tc::idesc_bf16_f32(kRows, NB);
    const uint32_t sbo = (uint32_t)(H >> 3) * 128;
    unsigned int *flag = p.flags + (d * gridDim.y + slice);

    constexpr int RPT = NB / 4;                 // batch rows per thread in the cell phase
    float c_state[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) c_state[j] = 0.f;

    const size_t p_row = (size_t)2 * 4 * H;     // elements per (b, t) in P
    const int gate_col = warp * H + u * kUnits + lane;          // this thread's gate row inside a (b,t,d) block

    long long *prof = (g_prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? g_prof : nullptr;
    long long tprev = clock64();
    for (int step = 0; step < T; ++step) {
        const int t = d ? (T - 1 - step) : step;
        const int t_prev = d ? (t + 1) : (t - 1);
        // ---- prefetch the input-projection terms for (gate = warp, unit = lane), all NB batch rows ----
        float pre[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int b = b0 + j;
            pre[j] = (b < B) ? __bfloat162float(p.P[((size_t)b * T + t) * p_row + (size_t)d * 4 * H + gate_col]) : 0.f;
        }
        float acc[NB];
        if (step > 0) {
            // ---- wait for the whole group to have published h_{t_prev}, then gather it from L2 ----
            if (tid == 0) {
                const unsigned int target = (unsigned int)G * (unsigned int)step;
                while (ld_acquire(flag) < target) { }
            }
            __syncthreads();
            PROF_MARK(0);                               // flag wait
            const int chunks = H >> 3;
            for (int i = tid; i < NB * chunks; i += kLstmThreads) {
                const int r = (i & 7) | ((i / (8 * chunks)) << 3);
                const int c = (i >> 3) % chunks;
                const int b = b0 + r;
                const bf16 *src = p.Y + ((size_t)(b < B ? b : 0) * T + t_prev) * (2 * H) + d * H + c * 8;
                tc::cp_async16(sH + tc::kmajor_off(r, c * 8, H), src, b < B ? 16u : 0u);
            }
            tc::cp_async_commit();
            tc::cp_async_wait<0>();
            tc::fence_proxy_async();
            tc::fence_before_sync();
            __syncthreads();
            PROF_MARK(1);                               // h gather
            if (tid == 0) {
                tc::fence_after_sync();
                const uint32_t a0 = tc::smem_u32(sW), h0 = tc::smem_u32(sH);
                for (int k = 0; k < H / 16; ++k)
                    tc::mma_bf16(tmem, tc::smem_desc(a0 + k * 256, 128, sbo), tc::smem_desc(h0 + k * 256, 128, sbo), idesc, k > 0);
                tc::mma_commit(&s_bar);
            }
            tc::mbar_wait(&s_bar, (step - 1) & 1);
            tc::fence_after_sync();
            PROF_MARK(2);                               // MMA issue + completion
#pragma unroll
            for (int c0 = 0; c0 < NB; c0 += 16) {
                uint32_t v[16];
                tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[c0 + j] = __uint_as_float(v[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < NB; ++j) acc[j] = 0.f;       // h_0 = 0
        }
        // ---- gate non-linearity (warp 0: i, 1: f, 2: g (tanh), 3: o) ----
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const float x = acc[j] + pre[j];
            const float a = (warp == 2) ? tanh_f(x) : sigmoid_f(x);
            s_act[(warp * NB + j) * 32 + lane] = a;
            if (p.save && b0 + j < B)
                p.P[((size_t)(b0 + j) * T + t) * p_row + (size_t)d * 4 * H + gate_col] = __float2bfloat16_rn(a);
        }
        tc::fence_before_sync();
        __syncthreads();
        PROF_MARK(3);                                   // tmem load + P add + activation (+ gate save)
        // ---- cell update: thread = (unit = lane, batch rows warp*RPT ...) ----
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int jj = warp * RPT + j;
            const int b = b0 + jj;
            const float gi = s_act[(0 * NB + jj) * 32 + lane], gf = s_act[(1 * NB + jj) * 32 + lane];
            const float gg = s_act[(2 * NB + jj) * 32 + lane], go = s_act[(3 * NB + jj) * 32 + lane];
            const float c = gf * c_state[j] + gi * gg;
            c_state[j] = c;
            const float h = go * tanh_f(c);
            if (b < B) {
                const size_t o = ((size_t)b * T + t) * (2 * H) + d * H + u * kUnits + lane;
                p.Y[o] = __float2bfloat16_rn(h);
                if (p.C) p.C[o] = c;
            }
        }
        __syncthreads();
        PROF_MARK(4);                                   // cell update + Y / C stores
        if (tid == 0) {
            __threadfence();
            atomicAdd(flag, 1u);
        }
        PROF_MARK(5);                                   // fence + publish
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, NB < 32 ? 32 : NB);
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

extern "C" {

// Debug: d_prof = 8 zeroed int64 cycle counters filled by CTA (0,0,0) of the next LSTM launches; NULL disables.
int mlvae_debug_set_profile_buffer(void *d_prof) {
    long long *ptr = (long long *)d_prof;
    MLVAE_CHECK_CUDA(cudaMemcpyToSymbol(g_prof, &ptr, sizeof(ptr)));
    return MLVAE_OK;
}

// Scratch: one zeroed uint32 per (direction, batch slice).
size_t mlvae_lstm_scratch_bytes(int B) { return sizeof(unsigned int) * 2 * (size_t)((B + 15) / 16) + 256; }

int mlvae_lstm_fwd(void *d_p, const void *d_whh, void *d_y, float *d_c, int B, int T, int H, int save_gates,
                   void *d_scratch, void *stream) {
    MLVAE_REQUIRE(d_p && d_whh && d_y && d_scratch, MLVAE_ERR_INVALID_ARG, "lstm_fwd: missing buffers");
    MLVAE_REQUIRE(B > 0 && T > 0 && H > 0, MLVAE_ERR_INVALID_ARG, "lstm_fwd: bad sizes");
    MLVAE_REQUIRE(H % 32 == 0 && H % 16 == 0 && H <= 704, MLVAE_ERR_UNSUPPORTED,
                  "lstm_fwd: hidden size must be a multiple of 32 and <= 704 (W_hh slice resident in shared memory), got %d", H);
    cudaStream_t st = (cudaStream_t)stream;
    const int G = H / kUnits;
    const int sms = sm_count();
    // batch rows per CTA: smallest of {16, 32, 64} that lets every CTA be co-resident (1 CTA / SM)
    int NB = 16;
    while (NB < 64 && (int64_t)G * ((B + NB - 1) / NB) * 2 > sms) NB *= 2;
    const int slices = (B + NB - 1) / NB;
    MLVAE_REQUIRE((int64_t)G * slices * 2 <= sms, MLVAE_ERR_UNSUPPORTED,
                  "lstm_fwd: batch %d x hidden %d needs %d co-resident CTAs (> %d SMs)", B, H, G * slices * 2, sms);
    const size_t smem = (size_t)kRows * H * 2 + (size_t)NB * H * 2 + (size_t)4 * NB * 32 * 4;
    MLVAE_REQUIRE(smem <= 227 * 1024, MLVAE_ERR_UNSUPPORTED, "lstm_fwd: %zu bytes of shared memory needed", smem);
    MLVAE_CHECK_CUDA(cudaMemsetAsync(d_scratch, 0, sizeof(unsigned int) * 2 * slices, st));
    LstmFwdParams prm{(bf16 *)d_p, (const bf16 *)d_whh, (bf16 *)d_y, d_c, (unsigned int *)d_scratch, B, T, H, save_gates};
    void *args[] = {&prm};
    dim3 grid(G, slices, 2), block(kLstmThreads);
    const void *fn = nullptr;
    if (NB == 16) fn = (const void *)lstm_fwd_kernel<16>;
    else if (NB == 32) fn = (const void *)lstm_fwd_kernel<32>;
    else fn = (const void *)lstm_fwd_kernel<64>;
    MLVAE_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MLVAE_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, smem, st));
    return MLVAE_OK;
}

}  // extern "C"
