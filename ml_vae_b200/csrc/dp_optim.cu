// Data-parallel optimiser step over NVLink PEER MEMORY: gradient reduce-scatter + global-norm clip + Adam + parameter
// all-gather as TWO kernels that read the peers' gradient arenas and write the peers' parameter arenas directly
// (replaces ncclAllReduce + mlvae_adam_clip_step of the single-GPU path when world > 1):
//   models/md_model.py:77-87   loss.backward(); check_gradients; optimizer.step(); zero_grad  -- under DDP the reference
//                              all-reduces every gradient and then runs the full Adam update on every rank.
// Here rank r owns shard r of the flat arena (train_step.FlatArena; n / world elements):
//   kernel 1  dp_reduce_kernel   barrier A (every rank's backward is done) -> g[shard] = sum over ranks (fixed rank order, P2P
//                                loads, or one multimem.ld_reduce through the NVSwitch) -> own gradient arena; sum of squares
//                                of the shard -> every rank's sync block; barrier-B signal
//   kernel 2  dp_adam_kernel     barrier B -> global norm from the W shard sums (same order on every rank) -> clip, Adam on
//                                the shard only (1/W of the optimiser work per rank), updated float32 parameters and their
//                                bf16 shadow stored to EVERY rank's arena (P2P stores or multimem.st), the whole local
//                                gradient arena zeroed; barrier C (all peers' parameter stores have landed here)
// Traffic per rank: (W-1)/W of the arena in (gradients), 1.5 (W-1)/W out (parameters + bf16) -- or 1/W in and 1.5/W out with
// multicast -- against 2 (W-1)/W each way for a ring all-reduce followed by a full-arena Adam on every rank.
//
// Barriers are epoch counters in the peers' sync blocks (st.release.sys / ld.acquire.sys, monotonic, never reset), so the two
// launches are CUDA-graph safe.  Every spin is bounded (kSpinTimeoutNs): a missing peer sets `error` instead of hanging
// the GPU.  Results are identical on every rank by construction (one owner per element) and deterministic (fixed orders).
// A non-finite loss on ANY rank (or a non-finite reduced gradient) skips the update on EVERY rank.
#include "common.cuh"

namespace mlvae {
namespace {

constexpr int kDpThreads = 256;
constexpr int kDpMaxWorld = 8;
constexpr int kDpMaxGrid = 148 * 8;
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000 * 1000 * 1000;

struct DpSync {
    // written by the peers (slot = writer's rank)
    uint32_t flag_a[kDpMaxWorld], flag_b[kDpMaxWorld], flag_c[kDpMaxWorld];
    float sumsq[kDpMaxWorld];
    // local
    uint32_t epoch, ticket1, ticket2, error;
    float step, last_norm, last_coef, pad;
    float partial[kDpMaxGrid];
};

struct DpArgs {
    float *grads[kDpMaxWorld];
    float *params[kDpMaxWorld];
    __nv_bfloat16 *p16[kDpMaxWorld];
    DpSync *sync[kDpMaxWorld];
    float *mc_grads, *mc_params;
    __nv_bfloat16 *mc_p16;
    float *m, *v;
    const float *loss;
    int64_t n4, lo4, hi4;            // float4 units: arena size, this rank's shard
    int world, rank;
    double lr, b1, b2;
    float gscale, eps, max_norm;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// peer data: never through the non-coherent path, never cached in L1
__device__ __forceinline__ float4 ld_sys_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_f4(float4 *p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_sys_u2(uint2 *p, uint2 v) { asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ float4 mc_ld_reduce_f4(const float4 *p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_f4(float4 *p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_st_bf16x4(uint2 *p, uint2 v) {
    asm volatile("multimem.st.relaxed.sys.global.v2.bf16x2 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// threads t < world wait until slot t of `flags` has reached `epoch`; the whole CTA leaves together
__device__ __forceinline__ void wait_flags(const uint32_t *flags, int world, uint32_t epoch, uint32_t *error) {
    if ((int)threadIdx.x < world) {
        const unsigned long long t0 = global_ns();
        while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
            if (global_ns() - t0 > kSpinTimeoutNs) { *error = 1u; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kDpThreads) dp_reduce_kernel(const DpArgs a) {
    DpSync *me = a.sync[a.rank];
    const uint32_t epoch = me->epoch + 1u;
    const int tid = threadIdx.x;
    // barrier A: "my backward is complete" (stream order) -> every peer; then wait for every peer's signal
    if (blockIdx.x == 0 && tid < a.world) st_release_sys(&a.sync[tid]->flag_a[a.rank], epoch);
    wait_flags(me->flag_a, a.world, epoch, &me->error);

    float acc = 0.f;
    float4 *own = reinterpret_cast<float4 *>(a.grads[a.rank]);
    const int64_t stride = (int64_t)gridDim.x * kDpThreads;
    for (int64_t i = a.lo4 + (int64_t)blockIdx.x * kDpThreads + tid; i < a.hi4; i += stride) {
        float4 s;
        if (a.mc_grads) {
            s = mc_ld_reduce_f4(reinterpret_cast<const float4 *>(a.mc_grads) + i);
        } else {
            float4 g[kDpMaxWorld];
#pragma unroll
            for (int r = 0; r < kDpMaxWorld; ++r)
                if (r < a.world) g[r] = ld_sys_f4(reinterpret_cast<const float4 *>(a.grads[r]) + i);
            s = g[0];
#pragma unroll
            for (int r = 1; r < kDpMaxWorld; ++r)
                if (r < a.world) { s.x += g[r].x; s.y += g[r].y; s.z += g[r].z; s.w += g[r].w; }
        }
        own[i] = s;
        const float x = s.x * a.gscale, y = s.y * a.gscale, z = s.z * a.gscale, w = s.w * a.gscale;
        acc += (x * x + y * y) + (z * z + w * w);
    }
    const float bs = block_sum(acc);
    __shared__ bool s_last;
    if (tid == 0) {
        me->partial[blockIdx.x] = bs;
        __threadfence();
        s_last = atomicAdd(&me->ticket1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    // last CTA: shard sum in CTA order, published with the barrier-B signal
    __threadfence();
    float t = 0.f;
    for (int i = tid; i < (int)gridDim.x; i += kDpThreads) t += reinterpret_cast<volatile float *>(me->partial)[i];
    float total = block_sum(t);
    __shared__ float s_total;
    if (tid == 0) {
        if (a.loss && !isfinite(*a.loss)) total = __int_as_float(0x7fc00000);      // this rank's loss is not finite: nobody updates
        s_total = total;
        me->ticket1 = 0u;
    }
    __syncthreads();
    if (tid < a.world) {
        reinterpret_cast<volatile float *>(a.sync[tid]->sumsq)[a.rank] = s_total;
        __threadfence_system();
        st_release_sys(&a.sync[tid]->flag_b[a.rank], epoch);
    }
}

__global__ void __launch_bounds__(kDpThreads) dp_adam_kernel(const DpArgs a) {
    DpSync *me = a.sync[a.rank];
    const uint32_t epoch = me->epoch + 1u;
    const int tid = threadIdx.x;
    wait_flags(me->flag_b, a.world, epoch, &me->error);      // every rank's shard sum is here; nobody reads my gradients any more
    __shared__ float s_coef, s_norm;
    __shared__ bool s_skip;
    if (tid == 0) {
        float tot = 0.f;
        for (int r = 0; r < a.world; ++r) tot += reinterpret_cast<volatile float *>(me->sumsq)[r];
        const float norm = sqrtf(tot);
        float coef = a.max_norm > 0.f ? a.max_norm / (norm + 1e-6f) : 1.f;
        coef = fminf(coef, 1.f);
        s_skip = !isfinite(tot);
        s_norm = norm;
        s_coef = coef;
    }
    __syncthreads();
    const bool skip = s_skip;
    const float coef = s_coef * a.gscale;
    const float step = me->step + 1.f;
    const double bc1 = 1.0 - pow(a.b1, (double)step), bc2 = 1.0 - pow(a.b2, (double)step);
    const float step_size = (float)(a.lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    const float b2 = (float)a.b2, omb1 = (float)(1.0 - a.b1), omb2 = (float)(1.0 - a.b2);
    float4 *g4 = reinterpret_cast<float4 *>(a.grads[a.rank]);
    float4 *p4 = reinterpret_cast<float4 *>(a.params[a.rank]);
    float4 *m4 = reinterpret_cast<float4 *>(a.m), *v4 = reinterpret_cast<float4 *>(a.v);
    const int64_t stride = (int64_t)gridDim.x * kDpThreads;
    for (int64_t i = (int64_t)blockIdx.x * kDpThreads + tid; i < a.n4; i += stride) {
        if (i >= a.lo4 && i < a.hi4 && !skip) {
            const float4 gw = g4[i];
            float4 pw = p4[i], mw = m4[i], vw = v4[i];
            const float gg[4] = {gw.x * coef, gw.y * coef, gw.z * coef, gw.w * coef};
            float pp[4] = {pw.x, pw.y, pw.z, pw.w}, mm[4] = {mw.x, mw.y, mw.z, mw.w}, vv[4] = {vw.x, vw.y, vw.z, vw.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                mm[e] = mm[e] + (gg[e] - mm[e]) * omb1;
                vv[e] = vv[e] * b2 + omb2 * gg[e] * gg[e];
                const float denom = sqrtf(vv[e]) / bc2_sqrt + a.eps;
                pp[e] = pp[e] - step_size * (mm[e] / denom);
            }
            pw = make_float4(pp[0], pp[1], pp[2], pp[3]);
            m4[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
            v4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
            const __nv_bfloat162 lo = __floats2bfloat162_rn(pw.x, pw.y), hi = __floats2bfloat162_rn(pw.z, pw.w);
            const uint2 pk = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
            if (a.mc_params) {
                mc_st_f4(reinterpret_cast<float4 *>(a.mc_params) + i, pw);
                if (a.mc_p16) mc_st_bf16x4(reinterpret_cast<uint2 *>(a.mc_p16) + i, pk);
            } else {
#pragma unroll
                for (int r = 0; r < kDpMaxWorld; ++r)
                    if (r < a.world) {
                        st_sys_f4(reinterpret_cast<float4 *>(a.params[r]) + i, pw);
                        if (a.p16[r]) st_sys_u2(reinterpret_cast<uint2 *>(a.p16[r]) + i, pk);
                    }
            }
        }
        g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);                                     // zero_grad (the whole local arena)
    }
    // barrier C: my parameter stores have landed everywhere -> signal; the step ends when every peer's have landed here
    __threadfence_system();
    __shared__ bool s_last;
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&me->ticket2, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid < a.world) st_release_sys(&a.sync[tid]->flag_c[a.rank], epoch);
    wait_flags(me->flag_c, a.world, epoch, &me->error);
    if (tid == 0) {
        if (!skip) me->step = step;
        me->last_norm = s_norm;
        me->last_coef = s_coef;
        me->ticket2 = 0u;
        __threadfence();
        me->epoch = epoch;
    }
}

int g_dp_max_ctas = 0;      // tests: several "ranks" share ONE GPU, their spinning grids must all be resident

}  // namespace
}  // namespace mlvae

using namespace mlvae;

extern "C" {

size_t mlvae_dp_sync_bytes(void) { return sizeof(DpSync); }

// Debug: cap the grids of the two kernels (0 = default).  Needed when several ranks are simulated on one device.
int mlvae_dp_debug_max_ctas(int n) { g_dp_max_ctas = n < 0 ? 0 : n; return MLVAE_OK; }

int mlvae_dp_adam_step(const mlvae_dp_adam_args *x, void *stream) {
    MLVAE_REQUIRE(x, MLVAE_ERR_INVALID_ARG, "dp_adam_step: null args");
    MLVAE_REQUIRE(x->world >= 1 && x->world <= kDpMaxWorld && x->rank >= 0 && x->rank < x->world, MLVAE_ERR_INVALID_ARG,
                  "dp_adam_step: world %d / rank %d (at most %d ranks of one node)", x->world, x->rank, kDpMaxWorld);
    MLVAE_REQUIRE(x->n > 0 && x->n % 4 == 0, MLVAE_ERR_INVALID_ARG, "dp_adam_step: arena size must be a positive multiple of 4 elements");
    MLVAE_REQUIRE(x->exp_avg && x->exp_avg_sq, MLVAE_ERR_INVALID_ARG, "dp_adam_step: missing moment buffers");
    DpArgs a{};
    for (int r = 0; r < x->world; ++r) {
        MLVAE_REQUIRE(x->grads[r] && x->params[r] && x->sync[r], MLVAE_ERR_INVALID_ARG, "dp_adam_step: missing peer pointer for rank %d", r);
        MLVAE_REQUIRE(((uintptr_t)x->grads[r] & 15) == 0 && ((uintptr_t)x->params[r] & 15) == 0 && ((uintptr_t)x->params_bf16[r] & 7) == 0 &&
                          ((uintptr_t)x->sync[r] & 15) == 0,
                      MLVAE_ERR_INVALID_ARG, "dp_adam_step: peer buffers must be 16-byte aligned");
        a.grads[r] = x->grads[r];
        a.params[r] = x->params[r];
        a.p16[r] = (__nv_bfloat16 *)x->params_bf16[r];
        a.sync[r] = (DpSync *)x->sync[r];
    }
    a.mc_grads = x->mc_grads;
    a.mc_params = x->mc_params;
    a.mc_p16 = (__nv_bfloat16 *)x->mc_params_bf16;
    MLVAE_REQUIRE((a.mc_grads == nullptr) == (a.mc_params == nullptr), MLVAE_ERR_INVALID_ARG, "dp_adam_step: multicast needs both the gradient and the parameter mapping");
    a.m = x->exp_avg;
    a.v = x->exp_avg_sq;
    a.loss = x->loss;
    a.n4 = x->n / 4;
    const int64_t per = (a.n4 + x->world - 1) / x->world;
    a.lo4 = per * x->rank < a.n4 ? per * x->rank : a.n4;
    a.hi4 = a.lo4 + per < a.n4 ? a.lo4 + per : a.n4;
    a.world = x->world;
    a.rank = x->rank;
    a.gscale = 1.f / (float)x->world;
    a.lr = x->lr; a.b1 = x->beta1; a.b2 = x->beta2; a.eps = (float)x->eps; a.max_norm = x->max_grad_norm;
    const int64_t cap_hw = (int64_t)sm_count() * 8 < kDpMaxGrid ? (int64_t)sm_count() * 8 : kDpMaxGrid;
    const int64_t cap = g_dp_max_ctas > 0 && g_dp_max_ctas < cap_hw ? g_dp_max_ctas : cap_hw;
    auto grid_for = [&](int64_t items) {
        const int64_t b = (items + kDpThreads - 1) / kDpThreads;
        return (int)(b < 1 ? 1 : b > cap ? cap : b);
    };
    cudaStream_t st = (cudaStream_t)stream;
    // a shard of a few hundred thousand float4: several loads per thread in flight hide the NVLink latency
    const int64_t cap1 = cap < (int64_t)sm_count() * 4 ? cap : (int64_t)sm_count() * 4;
    int g1 = grid_for(a.hi4 - a.lo4);
    if (g1 > cap1) g1 = (int)cap1;
    dp_reduce_kernel<<<g1, kDpThreads, 0, st>>>(a);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    dp_adam_kernel<<<grid_for(a.n4), kDpThreads, 0, st>>>(a);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

// {completed steps (epoch), Adam step count, last gradient norm, last clip coefficient, error flag} of a sync block -> host floats
int mlvae_dp_read_state(const void *d_sync, float out[5], void *stream) {
    MLVAE_REQUIRE(d_sync && out, MLVAE_ERR_INVALID_ARG, "dp_read_state: null");
    DpSync h;
    // only the header is needed; the struct is small (< 6 KB)
    MLVAE_CHECK_CUDA(cudaMemcpyAsync(&h, d_sync, offsetof(DpSync, partial), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MLVAE_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    out[0] = (float)h.epoch; out[1] = h.step; out[2] = h.last_norm; out[3] = h.last_coef; out[4] = (float)h.error;
    return MLVAE_OK;
}

// Restore the Adam step count of a sync block (checkpoint resume); every rank sets the same value before the next step.
int mlvae_dp_set_adam_step(void *d_sync, float step, void *stream) {
    MLVAE_REQUIRE(d_sync && step >= 0.f, MLVAE_ERR_INVALID_ARG, "dp_set_adam_step: bad arguments");
    MLVAE_CHECK_CUDA(cudaMemcpyAsync((char *)d_sync + offsetof(DpSync, step), &step, sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    MLVAE_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return MLVAE_OK;
}

}  // extern "C"
