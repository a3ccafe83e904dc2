// Joint boundary / mispronunciation decoder on the GPU (SURVEY.md section 8 f4 "later"; the reference runs it as a python
// triple loop per utterance under joblib):
//   utils/decode_utils.py:374-565   decode_plvl_md_lbl_seqs_full   (same DP: decode_plvl_md_lbl_seqs_full_non_par, :191-371)
// A Viterbi search over (phoneme l, frame t, beta = correct / mispronounced), then a backtrack that yields three INTEGER
// sequences (boundaries, frame-level labels, phoneme-level labels) which must equal the reference's bit for bit.  So the
// arithmetic is the reference's, not a convenient one: float64 sums of float32 terms, evaluated left to right exactly as
// decode_utils.py:452-500 writes them (explicit __dadd_rn / __dsub_rn / __dmul_rn: no contraction into FMAs), np.argmax's
// "first maximum wins" as strict comparisons in list order, the strict `>` of the final state choice (:508).  The two spots
// whose type depends on the numpy version (python float x np.float32, the all-float32 first cell; see oracle/decode_ref.py)
// follow `numpy2`.
//
// One CTA per utterance, thread = phoneme l (both beta states), one __syncthreads per frame: dp[l-1, t-1, :] comes from a
// double-buffered shared array, the emission gather log_p_yx[t, y_l, :] is prefetched a chunk of four frames ahead, the 2 x 2-bit back
// pointers of a (l, t) cell are one byte -- kept in shared memory when T x Lmax bytes fit (up to 226 KB: 20 s x 100 phonemes;
// the backtrack is a chain of T dependent reads: ~30 cycles each from shared memory, an L2 round trip each from global
// memory), in the caller's workspace otherwise.
#include "common.cuh"

namespace mlvae {
namespace {

constexpr size_t kDecSmemBudget = 226 * 1024;   // dp exchange + back pointers of one utterance stay on chip up to this many bytes
__host__ inline size_t dec_smem_bytes(int T, int Lmax) { return (size_t)4 * Lmax * sizeof(double) + (size_t)T * Lmax; }

struct DecodeParams {
    const float *log_p_yx, *log_p_b, *log_p_pi, *log_p_y;
    const int32_t *y, *feat_lens, *seq_lens;
    int B, T, N, Lmax;
    double weight;
    int numpy2;
    unsigned char *path_g;          // (B, T, Lmax) bytes or nullptr (shared-memory path)
    int32_t *boundary, *frames, *phones, *status;
};

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000ull); }

template <bool kPathInSmem, int kMaxThreads>
__global__ void __launch_bounds__(kMaxThreads) md_decode_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int i = blockIdx.x, l = threadIdx.x;
    const int T = p.T, N = p.N, Lmax = p.Lmax;
    const int Ti = p.feat_lens[i], Li = p.seq_lens[i];
    double *s_dp = reinterpret_cast<double *>(smem);                        // [2][Lmax][2]
    unsigned char *path = kPathInSmem ? smem + (size_t)4 * Lmax * sizeof(double) : p.path_g + (size_t)i * T * Lmax;
    __shared__ int s_bad;
    if (l == 0) s_bad = 0;
    __syncthreads();

    // outputs past the utterance's own lengths are -1
    for (int t = l; t < T; t += blockDim.x) {
        p.boundary[(size_t)i * T + t] = t < Ti ? 0 : -1;
        p.frames[(size_t)i * T + t] = -1;
    }
    for (int k = l; k < Lmax; k += blockDim.x) p.phones[(size_t)i * Lmax + k] = -1;
    const bool shape_ok = Li >= 1 && Ti >= 1 && Li <= Ti && Ti <= T && Li <= Lmax;       // the reference asserts l == t == 0 at the end
    const bool active = shape_ok && l < Li;
    int yl = 0;
    if (active) {
        yl = p.y[(size_t)i * Lmax + l];
        if (yl < 0 || yl >= N) { atomicExch(&s_bad, 2); yl = 0; }
    }
    __syncthreads();
    if (!shape_ok || s_bad) {
        if (l == 0) p.status[i] = shape_ok ? 2 : 1;
        return;
    }

    const float2 *lyx_row = reinterpret_cast<const float2 *>(p.log_p_yx) + (size_t)i * T * N + yl;      // + t * N
    const float2 *lb = reinterpret_cast<const float2 *>(p.log_p_b) + (size_t)i * T;
    const float2 *lpi = reinterpret_cast<const float2 *>(p.log_p_pi) + (size_t)i * T;
    const float wf = (float)p.weight;
    double dp0 = neg_inf(), dp1 = neg_inf();
    double ly0 = 0.0, ly1 = 0.0;
    if (active) {
        const float2 ly = __ldg(reinterpret_cast<const float2 *>(p.log_p_y) + yl);
        ly0 = (double)ly.x; ly1 = (double)ly.y;
        if (l == 0) {                                                       // decode_utils.py:452-453
            const float2 e = __ldg(lyx_row), pi0 = __ldg(lpi);
            if (p.numpy2) {
                dp0 = (double)__fsub_rn(__fadd_rn(__fmul_rn(wf, pi0.x), e.x), ly.x);
                dp1 = (double)__fsub_rn(__fadd_rn(__fmul_rn(wf, pi0.y), e.y), ly.y);
            } else {
                dp0 = __dsub_rn(__dadd_rn(__dmul_rn(p.weight, (double)pi0.x), (double)e.x), ly0);
                dp1 = __dsub_rn(__dadd_rn(__dmul_rn(p.weight, (double)pi0.y), (double)e.y), ly1);
            }
        }
        s_dp[(0 * Lmax + l) * 2 + 0] = dp0;
        s_dp[(0 * Lmax + l) * 2 + 1] = dp1;
    }
    // The per-frame inputs are prefetched one CHUNK of kPf frames ahead (an L2 / HBM round trip is longer than a DP step: ~150 cycles of
    // dependent float64 adds + one barrier), in registers with static indices.
    constexpr int kPf = 4;
    const float2 zero2 = make_float2(0.f, 0.f);
    float2 e_cur[kPf], b_cur[kPf], pi_cur[kPf];
#pragma unroll
    for (int k = 0; k < kPf; ++k) {
        const int t = 1 + k;
        const bool ld = active && t < Ti;
        e_cur[k] = ld ? __ldg(lyx_row + (size_t)t * N) : zero2;
        b_cur[k] = ld ? __ldg(lb + t) : zero2;
        pi_cur[k] = ld ? __ldg(lpi + t) : zero2;
    }
    __syncthreads();

    for (int t0 = 1; t0 < Ti; t0 += kPf) {
        float2 e_nxt[kPf], b_nxt[kPf], pi_nxt[kPf];
#pragma unroll
        for (int k = 0; k < kPf; ++k) {
            const int t = t0 + kPf + k;
            const bool ld = active && t < Ti;
            e_nxt[k] = ld ? __ldg(lyx_row + (size_t)t * N) : zero2;
            b_nxt[k] = ld ? __ldg(lb + t) : zero2;
            pi_nxt[k] = ld ? __ldg(lpi + t) : zero2;
        }
#pragma unroll
        for (int k = 0; k < kPf; ++k) {
            const int t = t0 + k;
            if (t >= Ti) break;                                  // uniform over the CTA
            if (active) {
                const float2 e = e_cur[k], bb = b_cur[k], pp = pi_cur[k];
                const double e0 = (double)e.x, e1 = (double)e.y, lb0 = (double)bb.x, lb1 = (double)bb.y;
                // hold: dp[l, t-1, s] + log_p_b[t, 0] + log_p_yx[t, y_l, s] - log_p_y[y_l, s]
                double v0 = __dsub_rn(__dadd_rn(__dadd_rn(dp0, lb0), e0), ly0);
                double v1 = __dsub_rn(__dadd_rn(__dadd_rn(dp1, lb0), e1), ly1);
                unsigned int c0 = 0, c1 = 0;
                if (l > 0) {
                    const double *prev = s_dp + ((size_t)((t - 1) & 1) * Lmax + (l - 1)) * 2;
                    const double q0 = prev[0], q1 = prev[1];
                    double w0, w1;
                    if (p.numpy2) { w0 = (double)__fmul_rn(wf, pp.x); w1 = (double)__fmul_rn(wf, pp.y); }
                    else { w0 = __dmul_rn(p.weight, (double)pp.x); w1 = __dmul_rn(p.weight, (double)pp.y); }
                    const double a0 = __dadd_rn(q0, lb1), a1 = __dadd_rn(q1, lb1);
                    // value_list = [hold, from_correct, from_incorrect]; np.argmax: the first maximum wins
                    const double fc0 = __dsub_rn(__dadd_rn(__dadd_rn(a0, w0), e0), ly0), fi0 = __dsub_rn(__dadd_rn(__dadd_rn(a1, w0), e0), ly0);
                    const double fc1 = __dsub_rn(__dadd_rn(__dadd_rn(a0, w1), e1), ly1), fi1 = __dsub_rn(__dadd_rn(__dadd_rn(a1, w1), e1), ly1);
                    if (fc0 > v0) { v0 = fc0; c0 = 1; }
                    if (fi0 > v0) { v0 = fi0; c0 = 2; }
                    if (fc1 > v1) { v1 = fc1; c1 = 1; }
                    if (fi1 > v1) { v1 = fi1; c1 = 2; }
                }
                dp0 = v0; dp1 = v1;
                double *cur = s_dp + ((size_t)(t & 1) * Lmax + l) * 2;
                cur[0] = v0; cur[1] = v1;
                path[(size_t)t * Lmax + l] = (unsigned char)(c0 | (c1 << 4));
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < kPf; ++k) { e_cur[k] = e_nxt[k]; b_cur[k] = b_nxt[k]; pi_cur[k] = pi_nxt[k]; }
    }

    // ---- backtrack (decode_utils.py:503-536): one thread, a chain of T_i dependent reads ----
    if (l != 0) return;
    int32_t *bnd = p.boundary + (size_t)i * T, *fr = p.frames + (size_t)i * T, *ph = p.phones + (size_t)i * Lmax;
    int ll = Li - 1, t = Ti - 1;
    const double *fin = s_dp + ((size_t)(t & 1) * Lmax + ll) * 2;
    int beta = fin[0] > fin[1] ? 0 : 1;
    fr[t] = beta;
    ph[ll] = beta;
    int lab = beta, status = 0;
    while (t > 0) {
        const unsigned int c = (path[(size_t)t * Lmax + ll] >> (4 * beta)) & 3u;
        if (c != 0) {
            if (ll == 0) { status = 3; break; }              // cannot happen for a feasible shape (dp[-1] does not exist)
            --ll;
            bnd[t] = 1;
            beta = c == 1 ? 0 : 1;
            lab = beta;
            ph[ll] = beta;
        }
        fr[t - 1] = lab;
        --t;
    }
    bnd[0] = 1;
    if (status == 0 && ll != 0) status = 3;                   // the reference: assert l == t == 0
    p.status[i] = status;
}

}  // namespace
}  // namespace mlvae

using namespace mlvae;

extern "C" {

// Bytes of global workspace mlvae_md_decode needs for the back pointers (0: they fit in shared memory).
size_t mlvae_md_decode_workspace_bytes(int B, int T, int Lmax) {
    if (B <= 0 || T <= 0 || Lmax <= 0) return 0;
    return dec_smem_bytes(T, Lmax) <= kDecSmemBudget ? 0 : (size_t)B * T * Lmax;
}

int mlvae_md_decode(const float *d_log_p_yx, const float *d_log_p_b, const float *d_log_p_pi, const float *d_log_p_y,
                    const int32_t *d_y, const int32_t *d_feat_lens, const int32_t *d_seq_lens, int B, int T, int N, int Lmax,
                    double weight, int numpy2, void *d_workspace, int32_t *d_boundary, int32_t *d_frames, int32_t *d_phones,
                    int32_t *d_status, void *stream) {
    MLVAE_REQUIRE(d_log_p_yx && d_log_p_b && d_log_p_pi && d_log_p_y && d_y && d_feat_lens && d_seq_lens, MLVAE_ERR_INVALID_ARG,
                  "md_decode: missing input buffers");
    MLVAE_REQUIRE(d_boundary && d_frames && d_phones && d_status, MLVAE_ERR_INVALID_ARG, "md_decode: missing output buffers");
    MLVAE_REQUIRE(B > 0 && T > 0 && N > 0 && Lmax > 0, MLVAE_ERR_INVALID_ARG, "md_decode: bad sizes B=%d T=%d N=%d Lmax=%d", B, T, N, Lmax);
    MLVAE_REQUIRE(Lmax <= 1024, MLVAE_ERR_UNSUPPORTED, "md_decode: at most 1024 canonical phonemes per utterance (one thread each), got %d", Lmax);
    MLVAE_REQUIRE((((uintptr_t)d_log_p_yx | (uintptr_t)d_log_p_b | (uintptr_t)d_log_p_pi | (uintptr_t)d_log_p_y) & 7) == 0, MLVAE_ERR_INVALID_ARG,
                  "md_decode: log-probability arrays must be 8-byte aligned (pairs are read as float2)");
    const size_t need = mlvae_md_decode_workspace_bytes(B, T, Lmax);
    MLVAE_REQUIRE(need == 0 || d_workspace, MLVAE_ERR_INVALID_ARG, "md_decode: %zu bytes of workspace required for T=%d Lmax=%d", need, T, Lmax);
    DecodeParams p{d_log_p_yx, d_log_p_b, d_log_p_pi, d_log_p_y, d_y, d_feat_lens, d_seq_lens, B, T, N, Lmax, weight, numpy2 != 0,
                   need ? (unsigned char *)d_workspace : nullptr, d_boundary, d_frames, d_phones, d_status};
    const int threads = ((Lmax + 31) / 32) * 32;
    const size_t dp_bytes = (size_t)4 * Lmax * sizeof(double);
    cudaStream_t st = (cudaStream_t)stream;
    MLVAE_CHECK_CUDA(cudaMemsetAsync(d_status, 0, (size_t)B * sizeof(int32_t), st));
    // up to 256 phonemes (every real utterance) the kernel is compiled without the 64-register cap of a 1024-thread block
    const bool small = threads <= 256;
    if (need == 0) {
        const size_t smem = dec_smem_bytes(T, Lmax);
        const void *fn = small ? (const void *)md_decode_kernel<true, 256> : (const void *)md_decode_kernel<true, 1024>;
        MLVAE_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDecSmemBudget));
        if (small) md_decode_kernel<true, 256><<<B, threads, smem, st>>>(p);
        else md_decode_kernel<true, 1024><<<B, threads, smem, st>>>(p);
    } else {
        if (small) md_decode_kernel<false, 256><<<B, threads, dp_bytes, st>>>(p);
        else md_decode_kernel<false, 1024><<<B, threads, dp_bytes, st>>>(p);
    }
    MLVAE_CHECK_CUDA(cudaGetLastError());
    return MLVAE_OK;
}

}  // extern "C"
