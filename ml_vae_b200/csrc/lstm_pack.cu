// Parameter plumbing of the persistent LSTM layer (csrc/lstm.cu) in two launches per layer and step instead of ~20
// torch ops (cat / cast / gather / add):
//   pack    float32 masters in torch's layout (weight_ih_l*, weight_hh_l*, bias_ih_l*, bias_hh_l* and their _reverse
//           twins, modules/decoder.py:14-15)  ->  bf16 W_ih with rows in the kernels' (direction, unit, gate) order,
//           bf16 W_hh (2, 4H, H), float32 bias = b_ih + b_hh in kernel order (added in the GEMM epilogue in float32)
//   bias    per-slice bias-gradient partials of the backward recurrence -> ACCUMULATED into b_ih / b_hh gradients
//   unpack  float32 weight gradients in kernel row order  ->  ACCUMULATED into the eight float32 gradient tensors
//           in torch's (direction, gate, unit) order (bias gradient added to both b_ih and b_hh)
#include "common.cuh"

namespace mlvae {
namespace {

using bf16 = __nv_bfloat16;
constexpr int kPackThreads = 256;

struct LstmMasters {
    const float *w_ih[2], *w_hh[2], *b_ih[2], *b_hh[2];
};
struct LstmGrads {
    float *w_ih[2], *w_hh[2], *b_ih[2], *b_hh[2];
};

// kernel row r = unit * 4 + gate  <->  torch row gate * H + unit
__device__ __forceinline__ int torch_row(int r, int H) { return (r & 3) * H + (r >> 2); }

__global__ void __launch_bounds__(kPackThreads) lstm_pack_kernel(LstmMasters m, int In, int H, bf16 *__restrict__ w_ih_p,
                                                                 bf16 *__restrict__ w_hh, float *__restrict__ bias_p) {
    const int H4 = 4 * H;
    const int64_t n_ih = (int64_t)2 * H4 * (In / 4), n_hh = (int64_t)2 * H4 * (H / 4), n_b = 2 * H4;
    for (int64_t i = (int64_t)blockIdx.x * kPackThreads + threadIdx.x; i < n_ih + n_hh + n_b; i += (int64_t)gridDim.x * kPackThreads) {
        if (i < n_ih) {                                   // 4 consecutive columns of one W_ih row (In % 4 == 0)
            const int q = In / 4;
            const int64_t row = i / q;
            const int c4 = (int)(i - row * q), d = (int)(row / H4), r = (int)(row - (int64_t)d * H4);
            const float4 v = __ldg(reinterpret_cast<const float4 *>(m.w_ih[d] + (size_t)torch_row(r, H) * In) + c4);
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            reinterpret_cast<uint2 *>(w_ih_p + (size_t)row * In)[c4] =
                make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
        } else if (i < n_ih + n_hh) {                     // W_hh keeps torch's row order (the kernels index it by gate * H + unit)
            const int64_t k = i - n_ih;
            const int q = H / 4;
            const int64_t row = k / q;
            const int c4 = (int)(k - row * q), d = (int)(row / H4), r = (int)(row - (int64_t)d * H4);
            const float4 v = __ldg(reinterpret_cast<const float4 *>(m.w_hh[d] + (size_t)r * H) + c4);
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            reinterpret_cast<uint2 *>(w_hh + (size_t)row * H)[c4] =
                make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
        } else {
            const int k = (int)(i - n_ih - n_hh), d = k / H4, r = k - d * H4, t = torch_row(r, H);
            bias_p[k] = m.b_ih[d][t] + m.b_hh[d][t];
        }
    }
}

__global__ void __launch_bounds__(kPackThreads) lstm_unpack_grads_kernel(const float *__restrict__ dw_ih_p, const float *__restrict__ dw_hh_p0,
                                                                         const float *__restrict__ dw_hh_p1, const float *__restrict__ db,
                                                                         int In, int H, LstmGrads g) {
    const int H4 = 4 * H;
    const int64_t n_ih = (int64_t)2 * H4 * (In / 4), n_hh = (int64_t)2 * H4 * (H / 4), n_b = 2 * H4;
    for (int64_t i = (int64_t)blockIdx.x * kPackThreads + threadIdx.x; i < n_ih + n_hh + n_b; i += (int64_t)gridDim.x * kPackThreads) {
        if (i < n_ih) {
            const int q = In / 4;
            const int64_t row = i / q;
            const int c4 = (int)(i - row * q), d = (int)(row / H4), r = (int)(row - (int64_t)d * H4);
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(dw_ih_p + (size_t)row * In) + c4);
            float4 *dst = reinterpret_cast<float4 *>(g.w_ih[d] + (size_t)torch_row(r, H) * In) + c4;
            float4 a = *dst;
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            *dst = a;
        } else if (i < n_ih + n_hh) {
            const int64_t k = i - n_ih;
            const int q = H / 4;
            const int64_t row = k / q;
            const int c4 = (int)(k - row * q), d = (int)(row / H4), r = (int)(row - (int64_t)d * H4);
            const float *src = d ? dw_hh_p1 : dw_hh_p0;
            float4 *dst = reinterpret_cast<float4 *>(g.w_hh[d] + (size_t)torch_row(r, H) * H) + c4;
            float4 a = *dst;
            if (src) {                                    // NULL: T == 1, no recurrent gradient
                const float4 v = __ldcs(reinterpret_cast<const float4 *>(src + (size_t)r * H) + c4);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            *dst = a;
        } else {                                          // db arrives in torch order already (reduced inside the backward kernel)
            const int k = (int)(i - n_ih - n_hh), d = k / H4, t = k - d * H4;
            const float v = db[k];
            g.b_ih[d][t] += v;
            g.b_hh[d][t] += v;
        }
    }
}

// db (torch order, (slices, 2, 4H)) summed over the slices in index order and added to the four bias gradients
__global__ void __launch_bounds__(kPackThreads) lstm_bias_grads_kernel(const float *__restrict__ db_part, int slices, int H, float *__restrict__ g_ih_f,
                                                                       float *__restrict__ g_hh_f, float *__restrict__ g_ih_r, float *__restrict__ g_hh_r) {
    const int H4 = 4 * H;
    for (int k = blockIdx.x * kPackThreads + threadIdx.x; k < 2 * H4; k += gridDim.x * kPackThreads) {
        float acc = 0.f;
        for (int s = 0; s < slices; ++s) acc += db_part[(size_t)s * 2 * H4 + k];
        const int d = k / H4, t = k - d * H4;
        if (d == 0) { g_ih_f[t] += acc; g_hh_f[t] += acc; } else { g_ih_r[t] += acc; g_hh_r[t] += acc; }
    }
}

int grid_for_items(int64_t n) {
    int64_t b = (n + kPackThreads - 1) / kPackThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(b < 1 ? 1 : b > cap ? cap : b);
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

extern "C" {

// masters[8] = {w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r} float32 device pointers (torch's order of
// nn.LSTM parameters for one layer).  In % 4 == 0, H % 4 == 0, all 16-byte aligned.
int mlvae_lstm_pack_weights(const float *const *masters, int In, int H, void *d_w_ih_p, void *d_w_hh, void *d_bias_p, void *stream) {
    MLVAE_REQUIRE(masters && d_w_ih_p && d_w_hh && d_bias_p, MLVAE_ERR_INVALID_ARG, "lstm_pack_weights: missing buffers");
    MLVAE_REQUIRE(In > 0 && H > 0 && In % 4 == 0 && H % 4 == 0, MLVAE_ERR_UNSUPPORTED, "lstm_pack_weights: In and H must be multiples of 4");
    LstmMasters m;
    for (int d = 0; d < 2; ++d) {
        m.w_ih[d] = masters[4 * d]; m.w_hh[d] = masters[4 * d + 1]; m.b_ih[d] = masters[4 * d + 2]; m.b_hh[d] = masters[4 * d + 3];
        MLVAE_REQUIRE(m.w_ih[d] && m.w_hh[d] && m.b_ih[d] && m.b_hh[d], MLVAE_ERR_INVALID_ARG, "lstm_pack_weights: NULL parameter");
        MLVAE_REQUIRE(((uintptr_t)m.w_ih[d] & 15) == 0 && ((uintptr_t)m.w_hh[d] & 15) == 0, MLVAE_ERR_INVALID_ARG,
                      "lstm_pack_weights: weights must be 16-byte aligned");
    }
    const int64_t n = (int64_t)8 * H * (In / 4) + (int64_t)8 * H * (H / 4) + 8 * H;
    lstm_pack_kernel<<<grid_for_items(n), kPackThreads, 0, (cudaStream_t)stream>>>(m, In, H, (bf16 *)d_w_ih_p, (bf16 *)d_w_hh, (float *)d_bias_p);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

// d_db_part (slices, 2, 4H) float32 (mlvae_lstm_bwd) -> summed over the slices (fixed order) and ACCUMULATED into the four bias
// gradients (bias_ih / bias_hh get the same gradient: the layer adds them, modules/decoder.py:14-15 via nn.LSTM).
int mlvae_lstm_bias_grads(const float *d_db_part, int slices, int H, float *d_g_ih_f, float *d_g_hh_f, float *d_g_ih_r, float *d_g_hh_r,
                          void *stream) {
    MLVAE_REQUIRE(d_db_part && d_g_ih_f && d_g_hh_f && d_g_ih_r && d_g_hh_r && slices > 0 && H > 0, MLVAE_ERR_INVALID_ARG, "lstm_bias_grads: bad arguments");
    lstm_bias_grads_kernel<<<grid_for_items(8 * H), kPackThreads, 0, (cudaStream_t)stream>>>(d_db_part, slices, H, d_g_ih_f, d_g_hh_f, d_g_ih_r, d_g_hh_r);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

// grads[8]: float32 gradient tensors in the order of `masters` above, ACCUMULATED into.  d_dw_ih_p (8H x In) and
// d_dw_hh_p0/1 (4H x H each, may be NULL) have rows in the kernels' (unit, gate) order; d_db (8H) is in torch order.
int mlvae_lstm_unpack_grads(const float *d_dw_ih_p, const float *d_dw_hh_p0, const float *d_dw_hh_p1, const float *d_db, int In, int H,
                            float *const *grads, void *stream) {
    MLVAE_REQUIRE(d_dw_ih_p && d_db && grads, MLVAE_ERR_INVALID_ARG, "lstm_unpack_grads: missing buffers");
    MLVAE_REQUIRE(In > 0 && H > 0 && In % 4 == 0 && H % 4 == 0, MLVAE_ERR_UNSUPPORTED, "lstm_unpack_grads: In and H must be multiples of 4");
    LstmGrads g;
    for (int d = 0; d < 2; ++d) {
        g.w_ih[d] = grads[4 * d]; g.w_hh[d] = grads[4 * d + 1]; g.b_ih[d] = grads[4 * d + 2]; g.b_hh[d] = grads[4 * d + 3];
        MLVAE_REQUIRE(g.w_ih[d] && g.w_hh[d] && g.b_ih[d] && g.b_hh[d], MLVAE_ERR_INVALID_ARG, "lstm_unpack_grads: NULL gradient");
        MLVAE_REQUIRE(((uintptr_t)g.w_ih[d] & 15) == 0 && ((uintptr_t)g.w_hh[d] & 15) == 0, MLVAE_ERR_INVALID_ARG,
                      "lstm_unpack_grads: gradients must be 16-byte aligned");
    }
    const int64_t n = (int64_t)8 * H * (In / 4) + (int64_t)8 * H * (H / 4) + 8 * H;
    lstm_unpack_grads_kernel<<<grid_for_items(n), kPackThreads, 0, (cudaStream_t)stream>>>(d_dw_ih_p, d_dw_hh_p0, d_dw_hh_p1, d_db, In, H, g);
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

}  // extern "C"
