// Inter-layer dropout of the decoder's stacked biLSTM (modules/decoder.py:14-15: nn.LSTM(..., dropout=rnn_dropout)
// applies dropout to the output of every layer but the last; models/test_vanilla_vae/model.yaml: dec_rnn_dropout 0.15).
//
// The mask is a counter-based Philox4x32-10 stream (philox.cuh) instead of torch's stateful generator, so the
// backward pass regenerates it (nothing stored), a captured CUDA graph draws a fresh mask on every replay (device step
// counter added to the offset) and the host oracle (oracle/philox_ref.py: dropout_keep_mask) reproduces it bit for bit:
//   element i -> Philox block q = i / 8 (counter = (q lo, q hi, offset lo, offset hi), key = seed), 16-bit lane
//                l = i % 8: word l / 2, low half first;  keep  <=>  u16 >= thresh,  thresh = round(p * 65536)
//   y = keep ? x * (1 / (1 - p)) : 0       (scale evaluated in float32 like torch's dropout)
// One Philox call per 16-byte bf16 vector (8 elements): the kernel stays HBM bound (2 * s bytes per element).
// The same kernel is its own backward (dx = dy * mask * scale).
#include "common.cuh"
#include "philox.cuh"

namespace mlvae {
namespace {

constexpr int kDropThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kDropThreads) dropout_kernel(const T *__restrict__ x, T *__restrict__ y, int64_t n, uint32_t thresh,
                                                               float scale, uint64_t seed, uint64_t offset,
                                                               const uint64_t *__restrict__ d_offset_add) {
    constexpr int V = Vec<T>::N;                 // 8 (bf16) or 4 (f32) elements per 16-byte vector
    const PhiloxKey key(seed);
    if (d_offset_add) offset += *d_offset_add;
    const int64_t nvec = n / V;
    for (int64_t v = (int64_t)blockIdx.x * kDropThreads + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * kDropThreads) {
        const int64_t i0 = v * V;
        const uint64_t q = (uint64_t)i0 >> 3;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)), key);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
        Vec<T> a;
        a.load_stream(x + i0);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int l = (int)(i0 & 7) + e;     // 0..7 (f32 vectors cover lanes 0-3 or 4-7 of the block)
            const uint32_t u = (w[l >> 1] >> ((l & 1) * 16)) & 0xffffu;
            a.v[e] = (u >= thresh) ? a.v[e] * scale : 0.f;
        }
        a.store(y + i0);
    }
    // tail (n % V elements), one thread
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = nvec * V; i < n; ++i) {
            const uint64_t q = (uint64_t)i >> 3;
            const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)), key);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
            const int l = (int)(i & 7);
            const uint32_t u = (w[l >> 1] >> ((l & 1) * 16)) & 0xffffu;
            y[i] = from_f32<T>((u >= thresh) ? to_f32<T>(x[i]) * scale : 0.f);
        }
    }
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

extern "C" int mlvae_dropout(const void *d_x, void *d_y, int64_t n, float p, uint64_t seed, uint64_t offset,
                             const uint64_t *d_offset_add, int dtype, void *stream) {
    MLVAE_REQUIRE(d_x && d_y, MLVAE_ERR_INVALID_ARG, "dropout: null buffer");
    MLVAE_REQUIRE(n >= 0 && p >= 0.f && p < 1.f, MLVAE_ERR_INVALID_ARG, "dropout: need n >= 0 and 0 <= p < 1 (got p = %f)", (double)p);
    MLVAE_REQUIRE(((uintptr_t)d_x & 15) == 0 && ((uintptr_t)d_y & 15) == 0, MLVAE_ERR_INVALID_ARG, "dropout: buffers must be 16-byte aligned");
    if (n == 0) return MLVAE_OK;
    const uint32_t thresh = (uint32_t)lrintf(p * 65536.f);
    const float scale = 1.f / (1.f - p);
    const int V = dtype == MLVAE_BF16 ? 8 : 4;
    int64_t blocks = (n / V + kDropThreads - 1) / kDropThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    const int grid = (int)(blocks < 1 ? 1 : blocks > cap ? cap : blocks);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MLVAE_BF16)
        dropout_kernel<__nv_bfloat16><<<grid, kDropThreads, 0, st>>>((const __nv_bfloat16 *)d_x, (__nv_bfloat16 *)d_y, n, thresh, scale, seed,
                                                                      offset, d_offset_add);
    else if (dtype == MLVAE_F32)
        dropout_kernel<float><<<grid, kDropThreads, 0, st>>>((const float *)d_x, (float *)d_y, n, thresh, scale, seed, offset, d_offset_add);
    else
        return fail(MLVAE_ERR_INVALID_ARG, "dropout: dtype must be MLVAE_F32 or MLVAE_BF16");
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}
