// Probe: all-gather of a small per-CTA payload inside a 16-CTA thread-block cluster through distributed shared memory,
// the exchange a cluster version of the persistent LSTM would do every timestep.  Modes:
//   0  cp.async.bulk shared::cta -> shared::cluster (one bulk copy per peer, completes tx on the peer's mbarrier)
//   1  st.async 16-byte remote stores with mbarrier complete_tx
// Reports cycles per step (CTA 0) with and without emulated work between steps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o dsmem_probe dsmem_probe.cu && ./dsmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

template <int G>
__global__ void __launch_bounds__(512, 1) k(int mode, int payload, int steps, int work, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[2];
    unsigned char *recv = smem;                              // [2 parity][G][payload]
    unsigned char *send = smem + 2 * G * payload;            // [payload]
    const int tid = threadIdx.x;
    const uint32_t me = cluster_rank();
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    cluster_sync();
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int step = 0; step < steps; ++step) {
        const int par = step & 1;
        if (tid == 0) mbar_expect(&full[par], (uint32_t)(G * payload));
        // produce my payload
        for (int i = tid * 16; i < payload; i += 512 * 16) *reinterpret_cast<uint4 *>(send + i) = make_uint4(acc, step, me, i);
        if (mode == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid < G) {                                   // one bulk copy per peer, issued by G different threads
                const uint32_t dst = mapa(smem_u32(recv + ((size_t)par * G + me) * payload), tid);
                const uint32_t bar = mapa(smem_u32(&full[par]), tid);
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "r"(smem_u32(send)), "r"(payload), "r"(bar) : "memory");
            }
        } else {
            __syncthreads();
            // G * payload / 16 remote 16-byte stores spread over the CTA
            const int per_peer = payload / 16;
            for (int i = tid; i < G * per_peer; i += 512) {
                const int peer = i / per_peer, w = i - peer * per_peer;
                const uint4 v = *reinterpret_cast<const uint4 *>(send + w * 16);
                const uint32_t dst = mapa(smem_u32(recv + ((size_t)par * G + me) * payload + w * 16), peer);
                const uint32_t bar = mapa(smem_u32(&full[par]), peer);
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                             ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(bar) : "memory");
            }
        }
        mbar_wait(&full[par], (step >> 1) & 1);
        // consume: touch one word per thread
        acc += *reinterpret_cast<const uint32_t *>(recv + (size_t)par * G * payload + (tid * 16) % (G * payload));
        __syncthreads();
        if (work > 0) { long long s = clock64(); while (clock64() - s < work) {} __syncthreads(); }
    }
    long long t1 = clock64();
    cluster_sync();
    if (blockIdx.x == 0 && tid == 0) { out[0] = t1 - t0; out[1] = acc; }
}

template <int G>
void run(int clusters, int mode, int payload, int work, long long *out) {
    const int steps = 2000;
    const size_t smem = (size_t)(2 * G + 1) * payload;
    cudaFuncSetAttribute(k<G>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G * clusters); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<G>, mode, payload, steps, work, out);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[2] = {0, 0}; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("cluster %2d x%d mode %d payload %4d B/CTA (%5d B gathered) work %4d: %7.0f cycles/step (%s / %s)\n", G, clusters, mode, payload,
           G * payload, work, (double)h[0] / steps, cudaGetErrorString(e), cudaGetErrorString(e2));
}

int main() {
    long long *out; cudaMalloc(&out, 64);
    for (int mode : {0, 1})
        for (int payload : {1024, 2048})
            for (int work : {0, 1400}) {
                run<16>(4, mode, payload, work, out);
                run<8>(8, mode, payload, work, out);
            }
    run<16>(1, 0, 2048, 0, out);
    run<16>(7, 0, 1024, 1400, out);
    return 0;
}
