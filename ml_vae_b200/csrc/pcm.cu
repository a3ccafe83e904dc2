// On-GPU collation of a ragged PCM batch.  The reference decodes every wav on the host
// (librosa.load -> float32, src/utils/data_io.py:189-196), pads on the host and caches pickled FEATURES
// (data_io.py:67-97).  Here the host ships the batch as ONE ragged blob of raw 16-bit (or float32) samples -- half
// (or less) of the bytes of the padded float32 batch -- and this kernel scales, zero-pads and lays it out as the
// (B, N) float32 matrix mlvae_fbank_fwd reads.  int16 -> float uses the libsndfile/librosa convention x / 32768.
#include "common.cuh"

namespace mlvae {
namespace {

constexpr int kPcmThreads = 256;

// one thread = 8 consecutive samples of one utterance; utterance starts are 8-sample aligned inside the blob
template <typename T>
__global__ void __launch_bounds__(kPcmThreads) pcm_unpack_kernel(const T *__restrict__ blob, const int64_t *__restrict__ offsets,
                                                                 const int *__restrict__ lens, int64_t n_max, float scale,
                                                                 float *__restrict__ out) {
    const int b = blockIdx.y;
    const int64_t i0 = ((int64_t)blockIdx.x * kPcmThreads + threadIdx.x) * 8;
    if (i0 >= n_max) return;
    const int64_t len = lens[b];
    const T *src = blob + offsets[b] + i0;
    float v[8];
    if (i0 + 8 <= len) {
        if constexpr (sizeof(T) == 2) {
            const uint4 raw = __ldcs(reinterpret_cast<const uint4 *>(src));
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                v[2 * e] = (float)(short)(w[e] & 0xffffu) * scale;
                v[2 * e + 1] = (float)(short)(w[e] >> 16) * scale;
            }
        } else {
            const float4 a = __ldcs(reinterpret_cast<const float4 *>(src)), c = __ldcs(reinterpret_cast<const float4 *>(src) + 1);
            v[0] = a.x * scale; v[1] = a.y * scale; v[2] = a.z * scale; v[3] = a.w * scale;
            v[4] = c.x * scale; v[5] = c.y * scale; v[6] = c.z * scale; v[7] = c.w * scale;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (i0 + e < len) ? (float)src[e] * scale : 0.f;
    }
    float *dst = out + (size_t)b * n_max + i0;
    if (i0 + 8 <= n_max && (n_max & 3) == 0) {
        reinterpret_cast<float4 *>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4 *>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (i0 + e < n_max) dst[e] = v[e];
    }
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

extern "C" int mlvae_pcm_unpack(const void *d_blob, int sample_dtype, const int64_t *d_offsets, const int *d_lens, int B,
                                int64_t n_max, float scale, float *d_out, void *stream) {
    MLVAE_REQUIRE(d_blob && d_offsets && d_lens && d_out, MLVAE_ERR_INVALID_ARG, "pcm_unpack: missing buffers");
    MLVAE_REQUIRE(B > 0 && B <= 65535 && n_max > 0, MLVAE_ERR_INVALID_ARG, "pcm_unpack: bad sizes (B=%d, n_max=%lld)", B, (long long)n_max);
    MLVAE_REQUIRE(sample_dtype == 0 || sample_dtype == 1, MLVAE_ERR_INVALID_ARG, "pcm_unpack: sample_dtype 0 = int16, 1 = float32");
    MLVAE_REQUIRE(((uintptr_t)d_blob & 15) == 0 && ((uintptr_t)d_out & 15) == 0, MLVAE_ERR_INVALID_ARG, "pcm_unpack: 16-byte aligned buffers");
    const dim3 grid((unsigned)((n_max + kPcmThreads * 8 - 1) / (kPcmThreads * 8)), (unsigned)B);
    if (sample_dtype == 0)
        pcm_unpack_kernel<short><<<grid, kPcmThreads, 0, (cudaStream_t)stream>>>((const short *)d_blob, d_offsets, d_lens, n_max, scale, d_out);
    else
        pcm_unpack_kernel<float><<<grid, kPcmThreads, 0, (cudaStream_t)stream>>>((const float *)d_blob, d_offsets, d_lens, n_max, scale, d_out);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}
