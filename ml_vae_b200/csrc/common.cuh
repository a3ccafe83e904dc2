// Shared device/host helpers for libmlvae_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/mlvae_b200.h"

namespace mlvae {

// ---------------------------------------------------------------- errors ----
inline char *err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
inline int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}
#define MLVAE_CHECK_CUDA(expr)                                                                 \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::mlvae::fail(MLVAE_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                 __FILE__, __LINE__);                                          \
    } while (0)
#define MLVAE_REQUIRE(cond, code, ...)                    \
    do {                                                  \
        if (!(cond)) return ::mlvae::fail(code, __VA_ARGS__); \
    } while (0)

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}

// ------------------------------------------------------------ reductions ----
constexpr int kMaxPartials = 4096;   // >= any grid the streaming kernels launch
struct ReduceScratch {
    unsigned int ticket;             // self-resetting arrival counter
    unsigned int pad[3];
    float partial[kMaxPartials];     // one slot per CTA, written then summed in index order
    float partial_rows[kMaxPartials];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum over the CTA; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ float block_sum(float v) {
    __shared__ float s_w[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) s_w[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (int)((blockDim.x + 31) >> 5)) ? s_w[lane] : 0.f;
        v = warp_sum(v);
    }
    return v;
}

// Frames of row b that the reference's mask keeps:
//   #{t in [0,T): float(t) < lens_b * T}   with the product rounded to float32
// (utils/data_utils.py:88 -> length_to_mask: arange(T, dtype=float32) < lens*T).
__device__ __forceinline__ float mask_threshold(float len_b, int T) { return __fmul_rn(len_b, (float)T); }
__device__ __forceinline__ int valid_frames(float len_b, int T) {
    const float x = mask_threshold(len_b, T);
    if (!(x > 0.f)) return 0;          // also NaN -> every compare false
    if (x >= (float)T) return T;
    return (int)ceilf(x);              // t < x  <=>  t <= ceil(x) - 1
}

// ---------------------------------------------------------- vector I/O ------
template <typename T> struct Vec;      // 16-byte vectors
template <> struct Vec<float> {
    static constexpr int N = 4;
    float v[4];
    __device__ __forceinline__ void load(const float *p) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void load_stream(const float *p) {
        const float4 t = __ldcs(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float *p) const {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    float v[8];
    __device__ __forceinline__ void unpack(const uint4 t) {
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ void load(const __nv_bfloat16 *p) { unpack(__ldg(reinterpret_cast<const uint4 *>(p))); }
    __device__ __forceinline__ void load_stream(const __nv_bfloat16 *p) { unpack(__ldcs(reinterpret_cast<const uint4 *>(p))); }
    __device__ __forceinline__ void store(__nv_bfloat16 *p) const {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        __stcs(reinterpret_cast<uint4 *>(p), make_uint4(w[0], w[1], w[2], w[3]));
    }
};

template <typename T> __device__ __forceinline__ float to_f32(T x);
template <> __device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

}  // namespace mlvae
