/*
 * mlvae_b200.h -- C ABI of libmlvae_b200.so: the B200 (sm_100a) kernels behind the
 * ML-VAE data-parallel training hot path.
 *
 * Boundary rules
 *   - extern "C", plain pointers and sizes only; no torch / C++ types.
 *   - every pointer named d_* is a DEVICE pointer on the current CUDA device;
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every entry point returns MLVAE_OK (0) or a negative mlvae_status; nothing
 *     throws across the ABI.  mlvae_last_error() gives the text for the calling thread.
 *   - no entry point allocates device memory behind the caller's back except
 *     mlvae_fbank_plan_create (small constant tables, freed by _destroy); scratch
 *     buffers are caller-owned and sized by the *_scratch_bytes() helpers.
 *   - launches are asynchronous on `stream`; the caller owns synchronisation.
 *   - there is no CPU fallback anywhere behind this interface.
 *
 * Each function names the reference interface it replaces (paths relative to
 * the reference checkout, weiwei-ww/ML-VAE).  INTEGRATION.md shows the ctypes
 * binding and the yaml change a reference maintainer would make.
 */
#ifndef MLVAE_B200_H_
#define MLVAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLVAE_ABI_VERSION 1

typedef enum mlvae_status {
    MLVAE_OK = 0,
    MLVAE_ERR_INVALID_ARG = -1, /* null pointer, negative size, misaligned buffer */
    MLVAE_ERR_UNSUPPORTED = -2, /* configuration outside what the kernels implement */
    MLVAE_ERR_CUDA = -3,        /* CUDA runtime / launch error (text in mlvae_last_error) */
    MLVAE_ERR_NO_DEVICE = -4    /* no sm_100 device visible */
} mlvae_status;

typedef enum mlvae_dtype {
    MLVAE_F32 = 0,
    MLVAE_BF16 = 1
} mlvae_dtype;

/* reduction argument of utils/data_utils.py:67 apply_lens_to_loss */
typedef enum mlvae_reduction {
    MLVAE_RED_MEAN = 0,      /* sum(loss*mask) / sum(mask)               -> out[1] */
    MLVAE_RED_BATCHMEAN = 1, /* sum(loss*mask) / B                       -> out[1] */
    MLVAE_RED_BATCH = 2      /* per row b: sum_b(loss*mask)/sum_b(mask)  -> out[B] */
} mlvae_reduction;

/* loss_type argument of modules/decoder.py:11 Decoder(...) */
typedef enum mlvae_recon_type {
    MLVAE_RECON_LIKELIHOOD = 0, /* decoder.py:40-43 Gaussian NLL, eps = 1e-5 */
    MLVAE_RECON_MSE = 1         /* decoder.py:48-49 */
} mlvae_recon_type;

int mlvae_abi_version(void);
const char *mlvae_last_error(void);
/* number of SMs / compute capability major*10+minor of the current device */
int mlvae_device_info(int *sm_count, int *cc);

/* Bytes of caller-owned scratch every *_fwd reduction below needs.  The buffer
 * must be zero-filled ONCE after allocation (the kernels leave it zeroed). */
size_t mlvae_reduce_scratch_bytes(void);

/* ------------------------------------------------------------------------- *
 * Counter-based eps (replaces torch.randn_like at modules/vanilla_vae.py:39).
 * Stream definition: oracle/philox_ref.py.  Element i of the stream uses
 * Philox4x32-10 counter (i/4, offset), key seed; the same (seed, offset) in
 * _fwd and _bwd regenerates the same eps without storing it.
 * ------------------------------------------------------------------------- */
int mlvae_philox_u32(uint64_t seed, uint64_t offset, int64_t n, uint32_t *d_out, void *stream);
int mlvae_philox_normal(uint64_t seed, uint64_t offset, int64_t n, void *d_out, int dtype, void *stream);
/* The float32 kernels draw 4 normals per Philox call from 32-bit uniforms; the bf16 kernels draw EIGHT per call from 16-bit
 * uniforms (block i/8, word j -> pair (2j, 2j+1): u1 = (lo16+1) 2^-16, theta = (hi16-32768) pi/32768; oracle/philox_ref.py
 * philox_normal_v2), which takes the bf16 reparameterisation kernel off the instruction-issue bound.  mlvae_philox_normal
 * materialises the stream of the kernels of `dtype`; _ex separates the storage type from the kernel type whose stream is
 * wanted (e.g. float32 values of the bf16 kernels' eps for a parity test). */
int mlvae_philox_normal_ex(uint64_t seed, uint64_t offset, int64_t n, void *d_out, int out_dtype, int kernel_dtype, void *stream);

/* ------------------------------------------------------------------------- *
 * Fused reparameterise + KL (+ length-masked mean).
 * Replaces VanillaVAE.reparameterize (modules/vanilla_vae.py:37-40),
 * VanillaVAE.compute_kld_loss (:42-45) and, when d_kl_out != NULL, the
 * apply_lens_to_loss(.., 'mean') over it (utils/data_utils.py:67-104).
 *
 *   z        = mu + exp(0.5*logvar) * eps
 *   kl       = -0.5 * (1 + logvar - mu^2 - exp(logvar))              (B,T,L)
 *   mask     = float(t) < d_lens[b] * T   (float32, no rounding)
 *   d_kl_out = { sum(kl*mask) / (count*L), sum(kl*mask), count*L }     float[3]
 *
 * d_mu, d_logvar, d_z, d_kl_elem: (B,T,L) contiguous, element type `dtype`.
 * d_eps: same shape/dtype, or NULL to draw from Philox(seed, offset + *d_offset_add).
 * d_offset_add: optional DEVICE uint64 added to `offset` (a device-resident step counter keeps a
 *   captured CUDA graph drawing fresh eps on every replay); NULL = 0.
 * d_kl_elem (unreduced KL, the reference module contract) may be NULL.
 * d_kl_out may be NULL (then d_lens and d_scratch may be NULL too).
 * ------------------------------------------------------------------------- */
int mlvae_reparam_kl_fwd(const void *d_mu, const void *d_logvar, const void *d_eps,
                         uint64_t seed, uint64_t offset, const uint64_t *d_offset_add, const float *d_lens,
                         int B, int T, int L, int dtype,
                         void *d_z, void *d_kl_elem, float *d_kl_out, void *d_scratch,
                         void *stream);

/* Same with mu / logvar rows ld_in elements apart (0 = L): the two halves of ONE stacked (B,T,2L) projection output
 * (modules/vanilla_vae.py:23-24 computed as one GEMM against [W_mu; W_logvar]) are read in place, d_logvar = d_mu + L. */
int mlvae_reparam_kl_fwd_strided(const void *d_mu, const void *d_logvar, int64_t ld_in, const void *d_eps,
                                 uint64_t seed, uint64_t offset, const uint64_t *d_offset_add, const float *d_lens,
                                 int B, int T, int L, int dtype,
                                 void *d_z, void *d_kl_elem, float *d_kl_out, void *d_scratch, void *stream);

/*   g_elem   = d_grad_kl_elem (may be NULL) + d_grad_kl_mean[0] * mask / (count*L) (may be NULL)
 *   grad_mu     = grad_z + g_elem * mu
 *   grad_logvar = grad_z * 0.5*exp(0.5*logvar)*eps + g_elem * 0.5*(exp(logvar) - 1)
 * d_grad_z may be NULL (treated as zero).  d_grad_kl_mean is a DEVICE float scalar. */
int mlvae_reparam_kl_bwd(const void *d_mu, const void *d_logvar, const void *d_eps,
                         uint64_t seed, uint64_t offset, const uint64_t *d_offset_add, const void *d_grad_z,
                         const void *d_grad_kl_elem, const float *d_grad_kl_mean,
                         const float *d_lens, int B, int T, int L, int dtype,
                         void *d_grad_mu, void *d_grad_logvar, void *stream);
/* Strided form: inputs as in mlvae_reparam_kl_fwd_strided, gradient rows ld_out elements apart (0 = L): both gradients land
 * in one stacked (B,T,2L) buffer that feeds the projection's backward GEMMs without a concatenation pass. */
int mlvae_reparam_kl_bwd_strided(const void *d_mu, const void *d_logvar, int64_t ld_in, const void *d_eps,
                                 uint64_t seed, uint64_t offset, const uint64_t *d_offset_add,
                                 const void *d_grad_z, const void *d_grad_kl_elem, const float *d_grad_kl_mean,
                                 const float *d_lens, int B, int T, int L, int dtype,
                                 void *d_grad_mu, void *d_grad_logvar, int64_t ld_out, void *stream);

/* ------------------------------------------------------------------------- *
 * Fused reconstruction loss (+ length-masked mean).
 * Replaces Decoder.compute_recon_loss (modules/decoder.py:37-53) and the
 * apply_lens_to_loss over it (models/test_vanilla_vae/model.py:50).
 *   likelihood: 0.5*(log(2*pi)_f32 + logvar + (target-mean)^2 / (exp(logvar) + 1e-5))
 *   mse:        (target-mean)^2            (d_logvar ignored, may be NULL)
 * d_out = { masked mean, masked sum, count*D }  float[3]; d_elem may be NULL.
 * ------------------------------------------------------------------------- */
int mlvae_recon_fwd(const void *d_mean, const void *d_logvar, const void *d_target,
                    const float *d_lens, int B, int T, int D, int dtype, int loss_type,
                    void *d_elem, float *d_out, void *d_scratch, void *stream);

/* g_elem as above.  d_grad_target may be NULL (the reference never needs it:
 * target is the un-learned feature tensor).  d_grad_logvar is ignored for mse. */
int mlvae_recon_bwd(const void *d_mean, const void *d_logvar, const void *d_target,
                    const void *d_grad_elem, const float *d_grad_mean_scalar,
                    const float *d_lens, int B, int T, int D, int dtype, int loss_type,
                    void *d_grad_mean, void *d_grad_logvar, void *d_grad_target, void *stream);

/* ------------------------------------------------------------------------- *
 * GMM-VAE / hierarchical-VAE family.
 * mlvae_gmm_reparam_kl_*: GMMVAE.reparameterize + compute_kld_loss against a LEARNED prior
 *   (modules/gmm_vae.py:51-67), unreduced, flat over n elements (n = B*T*N*L):
 *     z  = mu + exp(0.5*logvar) * eps
 *     kl = -0.5 * (1 + logvar - plogvar - (exp(logvar) + (mu - pmu)^2) / (exp(plogvar) + 1e-5))
 * mlvae_apply_weight_*: utils/data_utils.py:32-64 apply_weight(x (M,N,C), w (M,N)) -> (M,C), used 8x per
 *   HierarchicalVAE.forward (modules/h_vae.py:45-60); the backward gives both grad_x and grad_w (the
 *   straight-through Gumbel-softmax gradient flows through w).
 * ------------------------------------------------------------------------- */
int mlvae_gmm_reparam_kl_fwd(const void *d_mu, const void *d_logvar, const void *d_prior_mu, const void *d_prior_logvar,
                             const void *d_eps, uint64_t seed, uint64_t offset, const uint64_t *d_offset_add, int64_t n,
                             int dtype, void *d_z, void *d_kl_elem, void *stream);
int mlvae_gmm_reparam_kl_bwd(const void *d_mu, const void *d_logvar, const void *d_prior_mu, const void *d_prior_logvar,
                             const void *d_eps, uint64_t seed, uint64_t offset, const uint64_t *d_offset_add,
                             const void *d_grad_z, const void *d_grad_kl_elem, int64_t n, int dtype, void *d_grad_mu,
                             void *d_grad_logvar, void *d_grad_prior_mu, void *d_grad_prior_logvar, void *stream);
int mlvae_apply_weight_fwd(const void *d_x, const void *d_w, int64_t M, int N, int C, int dtype, void *d_out, void *stream);
int mlvae_apply_weight_bwd(const void *d_x, const void *d_w, const void *d_grad_out, int64_t M, int N, int C, int dtype,
                           void *d_grad_x, void *d_grad_w, void *stream);

/* ------------------------------------------------------------------------- *
 * Stand-alone length-masked reduction = utils/data_utils.py:67-104
 * apply_lens_to_loss(loss (B,T,C), lens (B,), reduction) for callers that keep
 * the reference's unreduced-loss module contract.
 * d_out: float[1] (mean, batchmean) or float[B] (batch).
 * ------------------------------------------------------------------------- */
int mlvae_masked_reduce_fwd(const void *d_loss, const float *d_lens, int B, int T, int C,
                            int dtype, int reduction, float *d_out, void *d_scratch, void *stream);
int mlvae_masked_reduce_bwd(const float *d_grad_out, const float *d_lens, int B, int T, int C,
                            int dtype, int reduction, void *d_grad_loss, void *stream);

/* ------------------------------------------------------------------------- *
 * Fused acoustic front-end = speechbrain.lobes.features.Fbank as declared at
 * config/run.yaml:39-44 and called at utils/data_io.py:197-201
 * (framing, Hamming window, 400-point real FFT, power, triangular mel bank,
 *  10*log10, per-utterance top_db=80 floor, optional delta / delta-delta,
 *  optional Kaldi-length truncation, zero padding past each utterance).
 * Only the reference's geometry is implemented: win_length == n_fft == 400.
 * ------------------------------------------------------------------------- */
typedef struct mlvae_fbank_plan mlvae_fbank_plan;

/* h_window: 400 host floats or NULL (periodic Hamming computed internally).
 * h_melmat: (n_fft/2+1) x n_mels row-major host floats or NULL (SpeechBrain
 * triangular construction computed internally in float32). */
int mlvae_fbank_plan_create(mlvae_fbank_plan **plan, int sample_rate, int hop_samples, int n_fft,
                            int n_mels, int deltas, const float *h_window, const float *h_melmat);
int mlvae_fbank_plan_destroy(mlvae_fbank_plan *plan);
/* frames emitted for an utterance of n samples: full = 1 + n/hop; kept =
 * min(full, (n + hop/2)/hop)  (data_io.py:198-201, data_io_utils.py:156) */
int mlvae_fbank_frames(const mlvae_fbank_plan *plan, int64_t n_samples, int truncate_kaldi);
int mlvae_fbank_feature_dim(const mlvae_fbank_plan *plan);
size_t mlvae_fbank_scratch_bytes(const mlvae_fbank_plan *plan, int B, int64_t n_max);

/* d_wav: (B, n_stride) float32, row b valid for d_wav_len[b] samples (int32
 * device array; NULL = every row is n_max samples).  d_out: (B, t_out, D) of
 * `out_dtype`, rows past each utterance's frame count are written as zeros.
 * d_out_frames: int32[B] frames written per utterance (may be NULL). */
int mlvae_fbank_fwd(const mlvae_fbank_plan *plan, const float *d_wav, const int32_t *d_wav_len,
                    int B, int64_t n_max, int64_t n_stride, int truncate_kaldi,
                    void *d_out, int out_dtype, int t_out, int32_t *d_out_frames,
                    void *d_scratch, void *stream);

/* ------------------------------------------------------------------------- *
 * Global input normalisation = speechbrain.processing.features.InputNormalization(norm_type='global')
 * (models/test_vanilla_vae/model.yaml:14-15, model.py:24-25; arithmetic SB-recall): per-utterance mean /
 * unbiased std over round(len*T) valid frames, batch average, running update with weight 1/(count+1).
 * State {count, pad[3], glob_mean[D], glob_std[D]} lives on the device (zero it once).
 * ------------------------------------------------------------------------- */
size_t mlvae_norm_state_bytes(int D);
size_t mlvae_norm_scratch_bytes(int B, int D);
int mlvae_global_norm(const float *d_x, const float *d_lens, int B, int T, int D, int training, int update_stats,
                      float *d_state, float *d_scratch, void *d_out, int out_dtype, void *stream);
/* Data-parallel statistics (optional): the training-mode call above in two halves with the CALLER's all-reduce(sum) of
 * d_avg {mean[D], std[D]} over the ranks in between, so that every rank keeps the running statistics of the GLOBAL batch (the
 * reference under DDP would keep per-process statistics, models/test_vanilla_vae/model.py:24-25).  avg_scale = 1 / world. */
int mlvae_global_norm_batch_avg(const float *d_x, const float *d_lens, int B, int T, int D, float *d_scratch, float *d_avg, void *stream);
int mlvae_global_norm_from_avg(const float *d_x, int B, int T, int D, const float *d_avg, float avg_scale, int update_stats,
                               float *d_state, void *d_out, int out_dtype, void *stream);

/* ------------------------------------------------------------------------- *
 * Joint boundary / mispronunciation decoder (SURVEY 8f-4 "later"): the dynamic programme + backtrack of
 * src/utils/decode_utils.py:440-548 (decode_plvl_md_lbl_seqs_full :374-565, same DP as _non_par :191-371), one CTA per
 * utterance.  Inputs are the log-probability arrays the reference's pre-computation builds (decode_utils.py:421-438, float32):
 *   d_log_p_yx (B, T, N, 2)  d_log_p_b (B, T, 2)  d_log_p_pi (B, T, 2)  d_log_p_y (N, 2)   finite or -inf, 8-byte aligned
 *   d_y (B, Lmax) canonical phoneme indices in [0, N);  d_feat_lens / d_seq_lens (B) ABSOLUTE lengths T_b <= T, 1 <= L_b <= min(Lmax, T_b)
 * Arithmetic is the reference's: float64 sums of the float32 terms in the order decode_utils.py:452-500 writes them, np.argmax's
 * first-maximum rule, strict `>` for the final state.  numpy2 != 0: `weight * log_p_pi` and the first cell are float32 (numpy >= 2
 * promotion, NEP 50); 0: float64 (numpy 1.x).  With weight == 1.0 only the first cell differs.
 * Outputs (int32, -1 past the utterance's own length): d_boundary (B, T) 1 at the first frame of every phoneme,
 * d_frames (B, T) frame-level labels (0 correct / 1 mispronounced), d_phones (B, Lmax) phoneme-level labels;
 * d_status (B): 0 ok, 1 infeasible lengths (the reference fails its `assert l == t == 0`), 2 phoneme index out of range, 3 broken path.
 * d_workspace: mlvae_md_decode_workspace_bytes(B, T, Lmax) bytes (0: back pointers fit in shared memory, pass NULL).  Lmax <= 1024.
 * ------------------------------------------------------------------------- */
size_t mlvae_md_decode_workspace_bytes(int B, int T, int Lmax);
int mlvae_md_decode(const float *d_log_p_yx, const float *d_log_p_b, const float *d_log_p_pi, const float *d_log_p_y,
                    const int32_t *d_y, const int32_t *d_feat_lens, const int32_t *d_seq_lens, int B, int T, int N, int Lmax,
                    double weight, int numpy2, void *d_workspace, int32_t *d_boundary, int32_t *d_frames, int32_t *d_phones,
                    int32_t *d_status, void *stream);

/* ------------------------------------------------------------------------- *
 * Ragged PCM batch -> padded float32 matrix (replaces the host-side decode + pad of
 * src/utils/data_io.py:189-196 and the pickled feature cache of data_io.py:67-97 on the training path).
 * d_blob: the batch's utterances back to back, each start aligned to 8 samples (16 bytes for int16);
 * sample_dtype 0 = int16, 1 = float32; d_offsets[B] (in samples), d_lens[B];
 * out (B, n_max) float32 = sample * scale (1/32768 for int16: the libsndfile / librosa convention), zeros past d_lens[b].
 * ------------------------------------------------------------------------- */
int mlvae_pcm_unpack(const void *d_blob, int sample_dtype, const int64_t *d_offsets, const int *d_lens, int B,
                     int64_t n_max, float scale, float *d_out, void *stream);

/* ------------------------------------------------------------------------- *
 * tcgen05 / TMEM dense projections (modules/fc_block.py:4-21 and the mean/log_var
 * heads of vanilla_vae.py:22-24 and decoder.py:24-25).
 * ------------------------------------------------------------------------- */
/* Test hook: D (128 x N, f32) = A (128 x K, bf16) * B (N x K, bf16)^T through one
 * tcgen05.mma tile; checks the descriptor / TMEM conventions of csrc/tc05.cuh. */
/* Y (M x N, bf16, row stride ldy) = act(X (M x K, bf16, row stride ldx) W (N x K, bf16)^T + bias (N, f32 or NULL)).
 * One Linear (+ LeakyReLU(0.01) when leaky != 0) of modules/fc_block.py:9-16 on tcgen05 with the accumulator in
 * TMEM.  N <= 256, K % 8 == 0, ldx % 8 == 0, X / W 16-byte aligned.  The input gradient of the same layer is the
 * same call with W^T. */
int mlvae_linear_fwd(const void *d_x, const void *d_w, const float *d_bias, void *d_y, int M, int N, int K,
                     int ldx, int ldy, int leaky, void *stream);
int mlvae_tc05_selftest(const void *d_a, const void *d_b, float *d_d, int N, int K, int a_in_tmem, void *stream);

/* Backward prologue of one Linear(+LeakyReLU) of modules/fc_block.py:9-16 (what autograd derives from it):
 *   g = dy * (y > 0 ? 1 : slope)   (d_y / d_g both NULL for a bare Linear),   db[N] = sum over the M rows of g (f32).
 * bf16 (M x N, row stride ld elements), N % 8 == 0, N <= 2048.  d_scratch: mlvae_dense_bwd_scratch_bytes(N) bytes,
 * zeroed ONCE by the caller (self-resetting).  Deterministic. */
size_t mlvae_dense_bwd_scratch_bytes(int N);
int mlvae_dense_bwd_prep(const void *d_dy, const void *d_y, void *d_g, float *d_db, int64_t M, int N, int64_t ld, float slope,
                         void *d_scratch, int accumulate /* db += instead of db = */, void *stream);

/* ------------------------------------------------------------------------- *
 * check_gradients + optimizer.step + zero_grad of MDModel.fit_batch (models/md_model.py:82-87) for torch.optim.Adam
 * (models/test_vanilla_vae/model.yaml:45-47) over ONE flat float32 parameter arena, two launches:
 *   g *= grad_scale (1 / world_size after the all-reduce);  g *= min(1, max_grad_norm / (||g||_2 + 1e-6))   [clip_grad_norm_]
 *   step += 1;  m += (g - m)(1 - beta1);  v = beta2 v + (1 - beta2) g g;
 *   p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)                                 [torch Adam]
 *   g = 0;  p_bf16 = bf16(p)   (d_params_bf16 may be NULL)
 * A non-finite *d_loss (device float, may be NULL) skips the update like check_gradients does (the gradients are still zeroed),
 * step unchanged.
 * d_state: mlvae_adam_state_bytes() bytes, zeroed once by the caller ({step, last norm, last clip coefficient, partials}).
 * max_grad_norm <= 0 disables clipping.  Deterministic.  All buffers 16-byte aligned.
 * ------------------------------------------------------------------------- */
size_t mlvae_adam_state_bytes(void);
int mlvae_adam_clip_step(float *d_params, float *d_grads, float *d_exp_avg, float *d_exp_avg_sq, void *d_params_bf16, int64_t n, float grad_scale,
                         double lr, double beta1, double beta2, double eps, float max_grad_norm, void *d_state, const float *d_loss, void *stream);

/* ------------------------------------------------------------------------- *
 * Data-parallel optimiser step over NVLink peer memory (csrc/dp_optim.cu): what DDP around the reference does after
 * loss.backward() -- all-reduce of every gradient, then check_gradients + Adam + zero_grad on EVERY rank
 * (models/md_model.py:77-87) -- as a gradient reduce-scatter, a sharded clip + Adam and a parameter all-gather in two
 * launches that load the peers' gradient arenas and store into the peers' parameter arenas directly:
 *   rank r owns elements [r * ceil(n/4/world) * 4, ...) of the flat arena;
 *   g[shard] = sum_ranks grads_r[shard] (fixed rank order);  norm = sqrt(sum over ranks of their shard's sum of squares of g / world);
 *   clip / Adam as mlvae_adam_clip_step with grad_scale = 1 / world on the shard;  updated float32 parameters and bf16 shadow
 *   stored into every rank's arrays;  the whole local gradient arena zeroed.
 * grads[r] / params[r] / params_bf16[r] / sync[r]: rank r's arrays as mapped into THIS process (CUDA IPC or VMM peer
 * mappings; entry [rank] is the local array).  sync[r]: mlvae_dp_sync_bytes() bytes, zeroed once before the first step, holds
 * the inter-rank epoch flags and the Adam step count.  mc_*: optional multicast (NVLS) mappings of the same arrays; when
 * given, the reduction is one multimem.ld_reduce and the parameter stores are multimem.st (set both or neither).
 * exp_avg / exp_avg_sq: local, only the shard is touched.  loss: local device float or NULL; a non-finite loss on any rank
 * (or a non-finite reduced gradient) skips the update on every rank.  Every rank must make the same calls in the same
 * order; the two launches are CUDA-graph capturable.  A peer that never arrives sets the error flag after 20 s instead of
 * hanging (mlvae_dp_read_state -> {epoch, adam step, last norm, last clip coefficient, error}).
 * ------------------------------------------------------------------------- */
#define MLVAE_DP_MAX_WORLD 8
typedef struct mlvae_dp_adam_args {
    int world, rank;
    float *grads[MLVAE_DP_MAX_WORLD];
    float *params[MLVAE_DP_MAX_WORLD];
    void *params_bf16[MLVAE_DP_MAX_WORLD];     /* all NULL: no bf16 shadow */
    void *sync[MLVAE_DP_MAX_WORLD];
    float *mc_grads, *mc_params;
    void *mc_params_bf16;
    float *exp_avg, *exp_avg_sq;
    int64_t n;
    double lr, beta1, beta2, eps;              /* doubles like torch.optim.Adam's: 1 - beta is rounded to float32 once */
    float max_grad_norm;
    const float *loss;
} mlvae_dp_adam_args;
size_t mlvae_dp_sync_bytes(void);
int mlvae_dp_adam_step(const mlvae_dp_adam_args *args, void *stream);
int mlvae_dp_read_state(const void *d_sync, float out[5], void *stream);
int mlvae_dp_set_adam_step(void *d_sync, float step, void *stream);   /* checkpoint resume: same value on every rank */
int mlvae_dp_debug_max_ctas(int n);          /* tests: cap both grids so that several ranks simulated on ONE device stay co-resident */

/* ------------------------------------------------------------------------- *
 * TMA-fed tcgen05 GEMM (csrc/gemm.cu) for the time-parallel matrix products of the step -- what torch dispatches
 * to cuBLAS for nn.LSTM / nn.Linear (modules/decoder.py:14-15,22 input projection x W_ih^T + b, its input / weight
 * gradients; modules/fc_block.py:9-16 weight gradients and wide forward layers):
 *     D_i[M, N] (+)= epilogue( A_i B_i ),   i < nprob problems of one shape in one launch, bf16 operands, f32 accumulation
 *   A: K-major   = row-major (M, K) with row stride lda, or
 *      MN-major  = row-major (K, M) with row stride lda (i.e. A^T without a transpose pass)
 *   B: K-major   = row-major (N, K) with row stride ldb (y = x W^T with W as nn.Linear stores it), or
 *      MN-major  = row-major (K, N) with row stride ldb
 *   batched reduction (both operands MN-major): D = sum_{b < kbatches} A_b B_b with K rows per batch, batch b at
 *      element offset b * a_batch_stride / b * b_batch_stride; rows past K of a batch read as zeros.  With row-shifted
 *      base pointers this is dW_hh = sum_b sum_t dA[b, t+1]^T h[b, t] in one call.
 *   epilogue: + bias_i[N] (f32, may be NULL), LeakyReLU(0.01) if leaky, the keep mask of mlvae_dropout over the flat
 *      (M, N) output if drop_p > 0 (needs ldd == N), bf16 or f32 output, accumulate (f32: D += ...), row_perm_H = H > 0:
 *      output row of GEMM row m is (m % 4) * H + m / 4 (the LSTM kernels' (unit, gate) row order -> torch's (gate, unit)).
 *   split_k > 1: the reduction is cut in split_k parts whose float32 partials go to ws
 *      (mlvae_gemm_workspace_bytes) and are added in split order by a second kernel (deterministic); f32 output only.
 *   bn: N tile (64, 128, 256; 0 = chosen from N).  N, lda, ldb %% 8 == 0; all pointers 16-byte aligned.
 *   max_ctas: cap of the persistent grid (0 = one CTA per SM): a GEMM that runs on a side stream BESIDE the persistent LSTM
 *      recurrence is confined to the SMs that launch leaves idle (lstm.py: the upper layer's dW_hh under the lower layer's backward).
 * ------------------------------------------------------------------------- */
#define MLVAE_GEMM_MAX_PROBLEMS 4
typedef struct mlvae_gemm_args {
    int nprob;
    const void *A[MLVAE_GEMM_MAX_PROBLEMS];
    const void *B[MLVAE_GEMM_MAX_PROBLEMS];
    void *D[MLVAE_GEMM_MAX_PROBLEMS];
    const float *bias[MLVAE_GEMM_MAX_PROBLEMS];
    int M, N, K, kbatches;
    int a_mn_major, b_mn_major;
    int64_t lda, ldb, a_batch_stride, b_batch_stride, ldd;
    int out_f32, accumulate, leaky, row_perm_H;
    int split_k;
    void *ws;
    float drop_p;
    uint64_t drop_seed, drop_offset;
    const void *drop_offset_add; /* device uint64[1] added to drop_offset, or NULL */
    int bn;
    int max_ctas;                /* > 0: at most this many (persistent) CTAs, e.g. the SMs a concurrently running cooperative kernel leaves idle */
} mlvae_gemm_args;
size_t mlvae_gemm_workspace_bytes(int nprob, int M, int N, int split_k);
int mlvae_gemm_bf16(const mlvae_gemm_args *args, void *stream);

/* ------------------------------------------------------------------------- *
 * Fused two-layer Linear / LeakyReLU(0.01) chain (csrc/mlp_chain.cu): modules/fc_block.py:4-21 as used by the encoder trunk
 * (modules/vanilla_vae.py:13-24, D -> 64 -> 64, both layers activated) and by the tails of the decoder heads
 * (modules/decoder.py:16-17,24-25, 64 -> 64 -> D, last layer bare; both heads = nprob 2 in one launch).
 *   forward   y_a = LeakyReLU(x W_a^T + b_a)  (saved for the backward pass if y_a != NULL),  y_b = act_b(y_a W_b^T + b_b)
 *   backward  g_b = g_out * act_b'(y_b);  dW_b += g_b^T y_a;  db_b += colsum g_b;  g_a = (g_b W_b) * LeakyReLU'(y_a);
 *             dW_a += g_a^T x;  db_a += colsum g_a;  dx = g_a W_a  (dx may be NULL)
 * A 128-row tile of the batch goes through both layers on chip (TMA -> tcgen05 -> TMEM -> shared-memory operand tile ->
 * tcgen05); the weight / bias gradients accumulate in tensor memory over the tiles of a CTA and are reduced in CTA order.
 * bf16 activations and weights (row-major W (N, K) as nn.Linear stores them, contiguous), float32 biases and gradients.
 * Widths: multiples of 16, K_A, N_A <= 112, N_B <= 128.  Every matrix 16-byte aligned with its leading dimension % 8 == 0.
 * ------------------------------------------------------------------------- */
typedef struct mlvae_chain_fwd_args {
    int nprob;              /* 1 or 2 independent chains of the same shape */
    const void *x[2];       /* (M, K_A) bf16, row stride ld_x */
    const void *w_a[2];     /* (N_A, K_A) bf16 */
    const void *w_b[2];     /* (N_B, N_A) bf16 */
    const float *bias_a[2], *bias_b[2];
    void *y_a[2];           /* (M, N_A) bf16, row stride ld_ya; NULL: not stored */
    void *y_b[2];           /* (M, N_B) bf16, row stride ld_yb */
    int M, K_A, N_A, N_B, act_b;
    int64_t ld_x, ld_ya, ld_yb;
} mlvae_chain_fwd_args;
int mlvae_mlp_chain_fwd(const mlvae_chain_fwd_args *args, void *stream);

typedef struct mlvae_chain_bwd_args {
    int nprob;
    const void *g_out[2];   /* (M, N_B) bf16 gradient of the chain output, row stride ld_g */
    const void *y_b[2];     /* (M, N_B) chain output, row stride ld_yb; only read when act_b */
    const void *y_a[2];     /* (M, N_A) hidden activation saved by the forward pass, row stride ld_ya */
    const void *x[2];       /* (M, K_A) chain input, row stride ld_x */
    const void *w_a[2], *w_b[2];
    float *dw_a[2], *db_a[2], *dw_b[2], *db_b[2];   /* float32, ACCUMULATED into */
    void *dx[2];            /* (M, K_A) bf16, row stride ld_dx, or NULL */
    int M, K_A, N_A, N_B, act_b;
    int64_t ld_g, ld_yb, ld_ya, ld_x, ld_dx;
    void *ws;               /* mlvae_mlp_chain_bwd_workspace_bytes(nprob, K_A, N_A) bytes */
} mlvae_chain_bwd_args;
size_t mlvae_mlp_chain_bwd_workspace_bytes(int nprob, int K_A, int N_A);
int mlvae_mlp_chain_bwd(const mlvae_chain_bwd_args *args, void *stream);

/* ------------------------------------------------------------------------- *
 * Inter-layer dropout of the stacked LSTM (modules/decoder.py:14-15: nn.LSTM(..., dropout=rnn_dropout),
 * models/test_vanilla_vae/model.yaml dec_rnn_dropout: 0.15), with a reproducible counter-based mask instead
 * of torch's stateful generator.  y = keep ? x / (1 - p) : 0; element i keeps iff the 16-bit lane i % 8
 * (word (i%8)/2, low half first) of Philox4x32-10(counter = (i/8, offset [+ *d_offset_add]), key = seed) is
 * >= round(p * 65536) (host restatement: oracle/philox_ref.py dropout_keep_mask).  The same call on dy is the
 * backward.  In-place (d_y == d_x) is allowed.  dtype MLVAE_F32 | MLVAE_BF16; buffers 16-byte aligned.
 * ------------------------------------------------------------------------- */
int mlvae_dropout(const void *d_x, void *d_y, int64_t n, float p, uint64_t seed, uint64_t offset,
                  const uint64_t *d_offset_add, int dtype, void *stream);

/* ------------------------------------------------------------------------- *
 * Persistent bidirectional LSTM recurrence (modules/decoder.py:14-15,22: nn.LSTM(batch_first,
 * bidirectional); one layer per call).  The input projection x W_ih^T + b_ih + b_hh for all
 * timesteps is a plain GEMM done by the caller into d_p; this entry point walks the T
 * recurrent steps of BOTH directions in one cooperative launch (tcgen05; every CTA's 128 x H
 * slice of W_hh stays resident in TENSOR MEMORY as the A operand).  bf16 only; H % 32 == 0, H <= 512.
 *   d_p   (B, T, 2, H, 4) bf16 gate pre-activations, UNIT-major with the four gates i,f,g,o of a unit
 *                             adjacent (permute the rows of W_ih / bias accordingly); when save_gates != 0
 *                             it is overwritten with the activated gates
 *   d_whh (2, 4H, H)    bf16  weight_hh_l{k}, weight_hh_l{k}_reverse
 *   d_y   (B, T, 2H)    bf16  output (forward direction in [:H], reverse in [H:])
 *   d_c   (B, T, 2H)    f32   cell states for the backward pass, or NULL
 *   d_scratch                 mlvae_lstm_scratch_bytes(B, H) bytes (zeroed by the call)
 * ------------------------------------------------------------------------- */
size_t mlvae_lstm_scratch_bytes(int B, int H);
/* debug: 8 zeroed int64 device counters receiving per-phase cycle totals of CTA 0; NULL disables */
int mlvae_debug_set_profile_buffer(void *d_prof);
/* debug: 4*T zeroed int32 device words receiving the four phase durations (cycles) of every step of one gate warp; NULL disables */
int mlvae_debug_set_trace_buffer(void *d_trace);
/* debug / tuning knobs: key 1 = MMA issuer warps of the LSTM kernels (1, 2, 4) */
int mlvae_debug_set_option(int key, int value);
int mlvae_lstm_fwd(void *d_p, const void *d_whh, void *d_y, float *d_c, int B, int T, int H,
                   int save_gates, void *d_scratch, void *stream);

/* Backward recurrence of the same layer.  d_gates: the (B,T,2,H,4) buffer mlvae_lstm_fwd filled with
 * activated gates (save_gates = 1); on return it holds the PRE-ACTIVATION gradients dA (bf16), from
 * which the caller forms dW_ih = dA^T x, dW_hh = dA^T h_prev, db = sum dA, dx = dA W_ih (plain GEMMs).
 * d_c: cell states from the forward pass; d_dy: (B,T,2H) bf16 gradient of the layer output.
 * d_bias_grad_part (may be NULL): (ceil(B/16), 2, 4H) float32, per-16-row-slice sums of dA over rows and time in
 * torch gate order; summing it over the first axis gives the bias gradient (deterministic). */
int mlvae_lstm_bwd(void *d_gates, const float *d_c, const void *d_dy, const void *d_whh, float *d_bias_grad_part,
                   int B, int T, int H, void *d_scratch, void *stream);

/* The same two kernels for ndir directions (2 = the calls above, 1 = a forward-only nn.LSTM(batch_first=True) layer: the main RNN
 * of the MD_VAE* recipes, src/models/MD_VAE/model.yaml:78-83, and src/modules/boundary_detector.py:19, phoneme_recognizer.py:13):
 * every "2" of the layouts above becomes ndir -- d_p / d_gates (B, T, ndir, H, 4), d_whh (ndir, 4H, H), d_y / d_c / d_dy (B, T, ndir*H),
 * d_bias_grad_part (ceil(B/16), ndir, 4H).  One cooperative launch of (H/32) * ceil(B/16) * ndir CTAs (at most one per SM). */
size_t mlvae_lstm_scratch_bytes_dirs(int B, int H, int ndir);
int mlvae_lstm_fwd_dirs(void *d_p, const void *d_whh, void *d_y, float *d_c, int B, int T, int H, int ndir,
                        int save_gates, void *d_scratch, void *stream);
int mlvae_lstm_bwd_dirs(void *d_gates, const float *d_c, const void *d_dy, const void *d_whh, float *d_bias_grad_part,
                        int B, int T, int H, int ndir, void *d_scratch, void *stream);

/* Parameter plumbing of one bidirectional layer (nn.LSTM's parameters, modules/decoder.py:14-15): masters[8] / grads[8] =
 * {weight_ih, weight_hh, bias_ih, bias_hh} of the forward direction then of the _reverse one, float32 device pointers in
 * torch's (gate, unit) row order.  pack: -> bf16 W_ih (8H x In) with rows in the kernels' (direction, unit, gate) order,
 * bf16 W_hh (2, 4H, H), FLOAT32 bias (8H) = b_ih + b_hh in kernel order (added by the GEMM epilogue).  unpack (only used
 * when the weight-gradient GEMMs cannot write torch's row order themselves): float32 gradients dW_ih (8H x In) and
 * dW_hh (4H x H per direction, NULL when T == 1) in kernel row order and db (8H, torch order, both biases) are
 * un-permuted and ACCUMULATED into grads[].  In % 4 == 0, H % 4 == 0, weights 16-byte aligned. */
int mlvae_lstm_pack_weights(const float *const *masters, int In, int H, void *d_w_ih_p, void *d_w_hh, void *d_bias_p, void *stream);
/* d_db_part (slices, 2, 4H): the per-slice bias-gradient partials mlvae_lstm_bwd wrote; summed in slice order and ACCUMULATED
 * into the four float32 bias gradients (bias_ih and bias_hh of a direction receive the same gradient). */
int mlvae_lstm_bias_grads(const float *d_db_part, int slices, int H, float *d_g_ih_f, float *d_g_hh_f, float *d_g_ih_r, float *d_g_hh_r,
                          void *stream);
int mlvae_lstm_unpack_grads(const float *d_dw_ih_p, const float *d_dw_hh_p0, const float *d_dw_hh_p1, const float *d_db, int In, int H,
                            float *const *grads, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MLVAE_B200_H_ */
