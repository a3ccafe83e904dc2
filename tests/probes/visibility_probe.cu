// Probe: how long does a 16-byte store by one SM take to become visible to a polling SM, as a function of the ADDRESS?
// (The backward LSTM's step time alternates with the parity buffer of its exchange words, and which buffer is the slow one follows
// the buffer's address: tests/probes/lstm_step_trace.py.)  Ping-pong between a producer CTA and a consumer CTA: the producer
// stores tag k at X, the consumer polls X, then stores the tag at a fixed acknowledge word Y which the producer polls; the
// producer's clock gives the round trip.  Y is fixed, so differences between X's are differences of the X leg.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o visibility_probe visibility_probe.cu && ./visibility_probe
#include <algorithm>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned ldv(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void stv(unsigned *p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned smid() { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }

// one thread per CTA takes part; CTAs other than `prod` / `cons` leave at once
__global__ void pingpong(unsigned *buf, const long long *offs, int n_offs, int reps, int prod, int cons, unsigned *ack, int *out, unsigned *smids) {
    if (threadIdx.x != 0) return;
    if ((int)blockIdx.x == prod) smids[0] = smid();
    if ((int)blockIdx.x == cons) smids[1] = smid();
    if ((int)blockIdx.x != prod && (int)blockIdx.x != cons) return;
    unsigned k = 0;
    for (int i = 0; i < n_offs; ++i) {
        unsigned *X = buf + offs[i] / 4;
        int best = 1 << 30;
        for (int r = 0; r < reps; ++r) {
            ++k;
            if ((int)blockIdx.x == prod) {
                const long long t0 = clock64();
                stv(X, k);
                while (ldv(ack) != k) {}
                const int dt = (int)(clock64() - t0);
                if (r > 0 && dt < best) best = dt;
            } else {
                while (ldv(X) != k) {}
                stv(ack, k);
            }
        }
        if ((int)blockIdx.x == prod) out[i] = best;
    }
}
// plain read latency of resident, unmodified data (dependent loads through L2) per address
__global__ void readlat(const unsigned *buf, const long long *offs, int n_offs, int *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int i = 0; i < n_offs; ++i) {
        const unsigned *X = buf + offs[i] / 4;
        unsigned v = ldv(X);
        int best = 1 << 30;
        for (int r = 0; r < 8; ++r) {
            const long long t0 = clock64();
            v += ldv(X + (v & 1u) * 0);
            const int dt = (int)(clock64() - t0 + (v == 0xffffffffu));
            if (dt < best) best = dt;
        }
        out[i] = best;
    }
}
int main() {
    const size_t bytes = 64ull << 20;
    unsigned *buf, *ack, *smids; int *out; long long *d_offs;
    cudaMalloc(&buf, bytes); cudaMalloc(&ack, 256); cudaMalloc(&smids, 8); cudaMemset(buf, 0, bytes); cudaMemset(ack, 0, 256);
    std::vector<long long> offs;
    for (long long o = 0; o < (16 << 20); o += 64 << 10) offs.push_back(o);            // 256 points, 64 KB apart
    const int n_coarse = (int)offs.size();
    for (long long o = 0; o < (256 << 10); o += 2 << 10) offs.push_back(o);            // 128 points, 2 KB apart
    const int n = (int)offs.size();
    cudaMalloc(&out, n * 4); cudaMalloc(&d_offs, n * 8); cudaMemcpy(d_offs, offs.data(), n * 8, cudaMemcpyHostToDevice);
    std::vector<int> h(n);
    readlat<<<1, 32>>>(buf, d_offs, n, out); cudaDeviceSynchronize(); cudaMemcpy(h.data(), out, n * 4, cudaMemcpyDeviceToHost);
    printf("read latency of clean resident data (cycles), 64 KB apart:");
    for (int i = 0; i < n_coarse; ++i) printf("%s%d", i % 32 ? " " : "\n  ", h[i]);
    printf("\n2 KB apart:");
    for (int i = n_coarse; i < n; ++i) printf("%s%d", (i - n_coarse) % 32 ? " " : "\n  ", h[i]);
    printf("\n");
    const int pairs[][2] = {{0, 1}, {0, 2}, {0, 75}, {0, 147}, {40, 100}};
    for (auto &pc : pairs) {
        cudaMemset(buf, 0, bytes); cudaMemset(ack, 0, 256);
        pingpong<<<148, 32>>>(buf, d_offs, n, 12, pc[0], pc[1], ack, out, smids);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h.data(), out, n * 4, cudaMemcpyDeviceToHost);
        unsigned sm[2]; cudaMemcpy(sm, smids, 8, cudaMemcpyDeviceToHost);
        std::vector<int> s(h.begin(), h.begin() + n_coarse); std::sort(s.begin(), s.end());
        printf("ping-pong CTA %d (SM %u) -> CTA %d (SM %u) [%s]: store-to-visible round trip, min of 11 (cycles); 64 KB apart: min %d median %d max %d", pc[0], sm[0], pc[1], sm[1],
               cudaGetErrorString(e), s.front(), s[n_coarse / 2], s.back());
        for (int i = 0; i < n_coarse; ++i) printf("%s%d", i % 32 ? " " : "\n  ", h[i]);
        printf("\n2 KB apart:");
        for (int i = n_coarse; i < n; ++i) printf("%s%d", (i - n_coarse) % 32 ? " " : "\n  ", h[i]);
        printf("\n");
    }
    return 0;
}
