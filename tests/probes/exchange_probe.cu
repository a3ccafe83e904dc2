// Probe: latency of the flag-in-data exchange through L2 that the persistent LSTM uses.
// Groups of G CTAs; every step each CTA publishes W words {payload, step} and waits until it has seen the
// words of all G members of its group for this step.  Reports cycles per step (CTA 0).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o exchange_probe exchange_probe.cu && ./exchange_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
__device__ __forceinline__ uint2 ldv(const uint2 *p) { uint2 v; asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void stv(uint2 *p, uint2 v) { asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory"); }
// words: [2 parity][groups][G producers][wpp words per producer]
__global__ void __launch_bounds__(512, 1) k(uint2 *ll, int G, int wpp, int steps, int work, long long *out, int threads_poll) {
    const int g = blockIdx.x / G, u = blockIdx.x % G, groups = gridDim.x / G, tid = threadIdx.x;
    const size_t gw = (size_t)G * wpp;
    long long t0 = clock64();
    unsigned acc = 0;
    for (int step = 1; step <= steps; ++step) {
        // publish my words for this step
        uint2 *dst = ll + ((size_t)(step & 1) * groups + g) * gw + (size_t)u * wpp;
        for (int i = tid; i < wpp; i += blockDim.x) stv(dst + i, make_uint2(acc + i, step));
        // gather all G * wpp words of the group: thread polls word tid, tid + 512, ...
        const uint2 *src = ll + ((size_t)(step & 1) * groups + g) * gw;
        if (tid < threads_poll)
            for (size_t i = tid; i < gw; i += threads_poll) {
                uint2 w = ldv(src + i);
                while (w.y != (unsigned)step) w = ldv(src + i);
                acc += w.x;
            }
        __syncthreads();
        if (work > 0) { long long s = clock64(); while (clock64() - s < work) {} __syncthreads(); }
    }
    if (blockIdx.x == 0 && tid == 0) { out[0] = clock64() - t0; out[1] = acc; }
}
// 16-byte words, tag bit replicated in bit 14 of every bf16 element; each thread polls its WPT words concurrently
__device__ __forceinline__ uint4 ldv4(const uint4 *p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void stv4(uint4 *p, uint4 v) { asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
template <int WPT>
__global__ void __launch_bounds__(512, 1) k16(uint4 *ll, int G, int steps, int work, long long *out) {
    const int g = blockIdx.x / G, u = blockIdx.x % G, groups = gridDim.x / G, tid = threadIdx.x;
    const int wpp = WPT * 512 / G;                       // words per producer
    const size_t gw = (size_t)WPT * 512;
    long long t0 = clock64();
    unsigned acc = 0;
    for (int step = 1; step <= steps; ++step) {
        const unsigned tagm = (((step + 1) >> 1) & 1) ? 0x40004000u : 0u;
        uint4 *dst = ll + ((size_t)(step & 1) * groups + g) * gw + (size_t)u * wpp;
        if (tid < wpp) { const unsigned v = (acc & 0x3fff3fffu) | tagm; stv4(dst + tid, make_uint4(v, v, v, v)); }
        const uint4 *src = ll + ((size_t)(step & 1) * groups + g) * gw;
        uint4 w[WPT];
#pragma unroll
        for (int n = 0; n < WPT; ++n) w[n] = ldv4(src + n * 512 + tid);
#pragma unroll
        for (int n = 0; n < WPT; ++n) {
            while (((w[n].x & 0x40004000u) != tagm) | ((w[n].y & 0x40004000u) != tagm) | ((w[n].z & 0x40004000u) != tagm) | ((w[n].w & 0x40004000u) != tagm))
                w[n] = ldv4(src + n * 512 + tid);
            acc += w[n].x + w[n].w;
        }
        __syncthreads();
        if (work > 0) { long long s = clock64(); while (clock64() - s < work) {} __syncthreads(); }
    }
    if (blockIdx.x == 0 && tid == 0) { out[0] = clock64() - t0; out[1] = acc; }
}
int main() {
    uint2 *ll; long long *out; cudaMalloc(&ll, 64 << 20); cudaMalloc(&out, 64);
    const int steps = 2000;
    struct Cfg { int ctas, G, wpp, work, tp; } cfgs[] = {
        {2, 2, 1, 0, 32}, {2, 2, 32, 0, 512}, {16, 16, 1, 0, 32}, {16, 16, 32, 0, 512}, {16, 16, 256, 0, 512}, {128, 16, 32, 0, 512},
        {128, 16, 256, 0, 512}, {128, 16, 256, 1400, 512}, {128, 16, 32, 1400, 512}, {128, 16, 128, 1400, 512}, {64, 16, 256, 1400, 512}, {16, 16, 256, 1400, 512}};
    for (auto c : cfgs) {
        cudaMemset(ll, 0, 64 << 20);
        void *args[] = {&ll, &c.G, &c.wpp, (void *)&steps, &c.work, &out, &c.tp};
        cudaError_t e = cudaLaunchCooperativeKernel((void *)k, dim3(c.ctas), dim3(512), args, 0, 0);
        cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("ctas %3d group %2d words/producer %3d work %4d: %6.0f cycles/step (%s)\n", c.ctas, c.G, c.wpp, c.work, (double)h[0] / steps,
               cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
    }
    for (int wpt : {1, 2, 4}) for (int work : {0, 1400}) for (int ctas : {16, 128}) {
        cudaMemset(ll, 0, 64 << 20);
        int G = 16;
        void *args[] = {&ll, &G, (void *)&steps, &work, &out};
        void *fn = wpt == 1 ? (void *)k16<1> : wpt == 2 ? (void *)k16<2> : (void *)k16<4>;
        cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(ctas), dim3(512), args, 0, 0);
        cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("16B words: ctas %3d words/thread %d work %4d: %6.0f cycles/step (%s)\n", ctas, wpt, work, (double)h[0] / steps,
               cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
    }
    return 0;
}
