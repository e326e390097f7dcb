/*
 * fdiff_b200.h — C ABI of the B200-native frequency-domain diffusion sampler.
 *
 * This is the drop-in boundary for the sampling hot path of JonathanCrabbe/FourierDiffusion ("fdiff").
 * The reference has no FFI of its own (it is pure Python on torch); each entry point below replaces the
 * reference Python call cited next to it (paths relative to the reference root, commit e60d532c).  The
 * Python host mirror in fourierdiffusion_b200/ binds these with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; fd_last_error() gives the message
 *     (thread-local, valid until the next failing call on the same thread);
 *   - `*_dev` pointers are device pointers on the handle's device, `*_host` pointers are host pointers;
 *     the library never frees caller memory and the caller keeps buffers alive until the stream reaches
 *     the end of the enqueued work;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous
 *     with respect to the host unless stated otherwise; calls on one handle are not re-entrant;
 *   - tensors are contiguous row-major fp32, series laid out (batch, max_len, n_channels) exactly like the
 *     reference's `DiffusableBatch.X` (src/fdiff/utils/dataclasses.py:7-18).
 */
#ifndef FDIFF_B200_H
#define FDIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FD_ABI_VERSION 1

/* score network family: src/fdiff/models/score_models.py:22 (ScoreModule), :249 (LSTMScoreModule), :169 (MLPScoreModule) */
enum { FD_MODEL_TRANSFORMER = 0, FD_MODEL_LSTM = 1, FD_MODEL_MLP = 2 };
/* scheduler family: src/fdiff/schedulers/sde.py:168 (VPScheduler), :90 (VEScheduler) */
enum { FD_SCHED_VP = 0, FD_SCHED_VE = 1 };
/* arithmetic of the score-network contractions.
 *   FD_MATH_FP32: fp32 FMA everywhere (generic kernels; any shape);
 *   FD_MATH_TF32: tensor-core path (tcgen05; operands rounded to 11 significant bits — TF32 for the attention projections and
 *                 Q·K^T, fp16 for the out-proj / FFN GEMMs and the softmax probabilities — fp32 accumulate) where the shape
 *                 has a specialised kernel, fp32 elsewhere.  The reference itself runs TF32 on CUDA (cmd/sample.py:23-24). */
enum { FD_MATH_FP32 = 0, FD_MATH_TF32 = 1 };

typedef struct fd_handle fd_handle;

typedef struct fd_config {
    int32_t struct_size;      /* = sizeof(fd_config), ABI guard */
    int32_t device;           /* CUDA device ordinal */
    int32_t model_kind;       /* FD_MODEL_* */
    int32_t max_len;          /* L   (ScoreModule.max_len,    score_models.py:38) */
    int32_t n_channels;       /* C   (ScoreModule.n_channels, score_models.py:39) */
    int32_t d_model;          /* D   (score_models.py:45) */
    int32_t n_head;           /* H   (transformer only; score_models.py:58) */
    int32_t num_layers;       /* score_models.py:60-62 / :276-286 / :205-210 */
    int32_t d_ff;             /* transformer: dim_feedforward (torch default 2048); MLP: d_mlp; LSTM: ignored */
    int32_t sched_kind;       /* FD_SCHED_* */
    double  sched_p0;         /* VP: beta_0 (sde.py:184)   VE: sigma_min (sde.py:105) */
    double  sched_p1;         /* VP: beta_1 (sde.py:185)   VE: sigma_max (sde.py:106) */
    int32_t fourier_noise_scaling; /* SDE.noise_scaling, sde.py:22 -> G of sde.py:42-60 */
    int32_t math_mode;        /* FD_MATH_* */
} fd_config;

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int         fd_abi_version(void);
const char *fd_last_error(void);
/* replaces: constructing DiffusionSampler around a ScoreModule, src/fdiff/sampling/sampler.py:12-22 */
int         fd_create(const fd_config *cfg, fd_handle **out);
int         fd_destroy(fd_handle *h);

/* ---- weights -------------------------------------------------------------------------------------------- */
/* Upload one tensor of the score module's state_dict by its reference key, e.g.
 * "backbone.layers.3.linear1.weight".  `data_host` holds `numel` fp32 values in the reference's layout.
 * The positional table must already be at the fixed point of nn.Embedding(max_norm) (transformer.py:13-15).
 * replaces: score_model.state_dict() being read by torch modules, score_models.py:53-62 */
int fd_set_weight(fd_handle *h, const char *name, const float *data_host, int64_t numel);
/* Validate that every tensor the configured model needs is present and build the packed device copies the
 * tensor-core kernels consume.  Synchronous. */
int fd_finalize_weights(fd_handle *h);

/* ---- per-phase entry points (parity tests; each is also a building block of fd_sample) ------------------- */
/* score = ScoreModule.forward / LSTMScoreModule.forward / MLPScoreModule.forward for a batch that shares one
 * diffusion time t.  replaces: score_models.py:67-94, :292-317, :215-246 */
int fd_score(fd_handle *h, const float *x_dev, float t, float *score_dev, int32_t batch, void *stream);
/* out = scheduler.step(model_output=score, timestep=t, sample=x) with the noise supplied by the caller.
 * `step_size` is the fp32 value of SDE.step_size (sde.py:64).  Bit-exact with the reference's CPU result.
 * replaces: VPScheduler.step sde.py:215-246, VEScheduler.step sde.py:129-165 */
int fd_step(fd_handle *h, const float *x_dev, const float *score_dev, const float *z_dev, double t, float step_size,
            float *out_dev, int32_t batch, void *stream);
/* out = SDE.prior_sampling from supplied standard-normal draws z (G ⊙ z, VE: × sigma_max).
 * replaces: sde.py:79-87, :125-127 (via sampler.py:111-122) */
int fd_prior(fd_handle *h, const float *z_dev, float *out_dev, int32_t batch, void *stream);
/* Transformer only: h <- LayerNorm2(h + linear2(relu(linear1(h)))) of encoder layer `layer`, in place on (n_tokens, d_model)
 * row-major activations — the FFN half of nn.TransformerEncoderLayer (score_models.py:57-62), exposed so the fused tensor-core
 * kernel can be checked in isolation. */
int fd_ffn_block(fd_handle *h, int32_t layer, float *h_dev, int32_t n_tokens, void *stream);
/* Transformer only: h <- LayerNorm1(h + out_proj(MHA(h))) of encoder layer `layer`, in place on (batch, max_len, d_model)
 * activations — the self-attention half of nn.TransformerEncoderLayer (score_models.py:57-62). */
int fd_attention_block(fd_handle *h, int32_t layer, float *h_dev, int32_t batch, void *stream);
/* Transformer only: h <- backbone(h), every encoder layer in place on (batch, max_len, d_model) activations — the
 * `self.backbone(X)` of score_models.py:87.  On the tensor-core path this is ONE launch of the persistent encoder-stack kernel
 * (csrc/fd_step.cu) unless option "persistent_stack" is 0. */
int fd_encoder_stack(fd_handle *h, float *h_dev, int32_t batch, void *stream);
/* Fill out_dev with the library's own counter-based standard normals (Philox4x32-10 + Box-Muller) for
 * `batch` series starting at global series index `first_series`; `draw` 0 is the prior draw, draw i+1 the noise of
 * diffusion step i.  Results do not depend on how series are sharded over GPUs.  (No reference equivalent: the
 * reference draws from torch's global generator, sde.py:85,238.) */
int fd_normal(fd_handle *h, uint64_t seed, uint64_t first_series, uint32_t draw, float *out_dev, int32_t batch, void *stream);

/* ---- evaluation loss (forward only) ------------------------------------------------------------------------ */
/* fd_score with one diffusion time PER SERIES (t_dev: batch fp32 values on the device) — ScoreModule.forward on a training / validation
 * style batch whose `timesteps` differ (score_models.py:67-94 with batch.timesteps of losses.py:58-62).  Same kernels as fd_score; the
 * time embedding becomes a (batch, d_model) table. */
int fd_score_t(fd_handle *h, const float *x_dev, const float *t_dev, float *score_dev, int32_t batch, void *stream);
/* out = mean(x0, t) + diag(std(t)) z: the forward perturbation of the SDE at per-series times with caller-supplied standard normals z —
 * scheduler.marginal_prob + scheduler.add_noise as the loss uses them (losses.py:67-84; sde.py:66-77, :108-123 VE, :187-210 VP).
 * std_scalar_dev (batch, may be NULL) receives the per-series scalar s_b with std_{b,l} = s_b * G_l.  Needs only the scheduler part of
 * the handle ("noise_scheduler.G"), no score-network weights. */
int fd_perturb(fd_handle *h, const float *x0_dev, const float *t_dev, const float *z_dev, float *out_dev, float *std_scalar_dev, int32_t batch,
               void *stream);
/* The denoising score-matching loss of one batch, forward only: perturb (fd_perturb), score (fd_score_t), then
 * loss_b = w_b * reduce((score + z / std)^2) with w_b = 1 / sum_l std_{b,l}^-2, or reduce((std * (score + z / std))^2) with
 * likelihood_weighting; reduce = mean over (L, C) if reduce_mean else 0.5 * sum; *loss_dev = mean_b loss_b.  losses_dev (batch) may be
 * NULL.  This is what ScoreModule.validation_step evaluates; the training step additionally needs dropout and the backward pass, which
 * this library does not provide.  replaces: get_sde_loss_fn(train=False), src/fdiff/utils/losses.py:39-125 (score_models.py:110-113) */
int fd_sde_loss(fd_handle *h, const float *x0_dev, const float *t_dev, const float *z_dev, int32_t likelihood_weighting, int32_t reduce_mean,
                float *losses_dev, float *loss_dev, int32_t batch, void *stream);

/* ---- the hot loop --------------------------------------------------------------------------------------- */
/* One batch of DiffusionSampler.sample: prior, then n_run reverse-diffusion steps on the time grid
 * `timesteps_host[0..n_run)` (fp32 values of SDE.timesteps, sde.py:63) with constant `step_size`.
 *   prior_z_dev : (batch, L, C) standard normals for the prior, or NULL -> Philox(seed, first_series, draw 0)
 *   noise_dev   : (n_run, batch, L, C) per-step normals,        or NULL -> Philox(seed, first_series, draw i+1)
 *   out_dev     : (batch, L, C) final sample, model domain (no de-standardise, no idft), like sampler.py:104.
 * Asynchronous on `stream`.  replaces: sampler.py:80-104 (inner loop of DiffusionSampler.sample) */
int fd_sample(fd_handle *h, int32_t batch, int32_t n_run, const float *timesteps_host, float step_size, uint64_t seed,
              uint64_t first_series, const float *prior_z_dev, const float *noise_dev, float *out_dev, void *stream);
/* Same, end to end with HOST buffers: noise (if given) is copied host->device, the result device->host
 * (the `X.cpu()` of sampler.py:107) and the call returns when out_host is complete.  prior_z_host / noise_host may be NULL. */
int fd_sample_host(fd_handle *h, int32_t batch, int32_t n_run, const float *timesteps_host, float step_size, uint64_t seed,
                   uint64_t first_series, const float *prior_z_host, const float *noise_host, float *out_host, void *stream);

/* ---- Fourier utilities (stateless) ---------------------------------------------------------------------- */
/* out = dft(x): ortho rFFT along dim 1 of (batch, L, C), packed real [Re X_0..X_{L/2} | Im X_1..X_{ceil(L/2)-1}].
 * replaces: src/fdiff/utils/fourier.py:8-45 */
int fd_dft(const float *x_dev, float *out_dev, int32_t batch, int32_t max_len, int32_t n_channels, int32_t device, void *stream);
/* out = idft(x * std + mean): the inverse of fd_dft with the de-standardisation of cmd/sample.py:76-78 fused in front
 * (mean_dev/std_dev: (L, C) or both NULL).  replaces: fourier.py:48-87 (+ cmd/sample.py:76-82) */
int fd_idft(const float *x_dev, float *out_dev, int32_t batch, int32_t max_len, int32_t n_channels, const float *mean_dev,
            const float *std_dev, int32_t device, void *stream);

/* out = spectral_density(x): |X_k|^2 of the ortho rFFT for k = 0 .. L/2, shape (batch, L/2 + 1, C) — the metrics front-end that
 * follows the sampler in cmd/sample.py:85 (metrics.py:76-84).  apply_dft = 0: x is already the packed spectrum (fd_dft output);
 * apply_dft = 1: x is a time series and `scratch_dev` a (batch, L, C) buffer for its spectrum.
 * replaces: src/fdiff/utils/fourier.py:90-124 */
int fd_spectral_density(const float *x_dev, float *out_dev, float *scratch_dev, int32_t batch, int32_t max_len, int32_t n_channels,
                        int32_t apply_dft, int32_t device, void *stream);

/* ---- evaluation metrics (stateless) ---------------------------------------------------------------------- */
/* Wasserstein-2 distances between two sample sets along n_dirs one-dimensional projections — the metric stage behind the sampler
 * (cmd/sample.py:85 -> metrics.py:102-199): out[k] = sqrt(emd2_1d(x . dir_k, y . dir_k)) with uniform weights (POT's ot.emd2_1d,
 * squared-Euclidean ground cost), any n and m.
 *   x_dev (n, d), y_dev (m, d): fp32 row-major flattened samples (check_flat_array, utils/tensors.py:5-24)
 *   dirs_dev (n_dirs, d) fp64 unit vectors (sliced distances), or NULL with n_dirs == d for the marginal distances (standard basis)
 *   standardise != 0: both projections are divided by the standard deviation of the x projection (normalisation = 'standardise')
 *   out_dev: n_dirs fp64 distances.  Asynchronous on `stream`.
 * replaces: WassersteinDistances.sliced_distances / marginal_distances, src/fdiff/utils/wasserstein.py:95-199 */
int fd_wasserstein(const float *x_dev, const float *y_dev, const double *dirs_dev, int32_t n, int32_t m, int32_t d, int32_t n_dirs,
                   int32_t standardise, double *out_dev, int32_t device, void *stream);

/* ---- data-set statistics (stateless) --------------------------------------------------------------------- */
/* Per-feature mean and unbiased standard deviation over the n series of x_dev (n, n_features) — n_features = L * C of a (n, L, C)
 * tensor; the (L, C) statistics that standardise the (DFT'd) training set and de-standardise the samples (cmd/sample.py:76-78).
 * replaces: DiffusionDataset.__init__ `X_ref.mean(dim=0)`, `X_ref.std(dim=0)`, src/fdiff/dataloaders/datamodules.py:52-53 */
int fd_feature_stats(const float *x_dev, float *mean_dev, float *std_dev, int64_t n, int32_t n_features, int32_t device, void *stream);
/* out = (x - mean) / std (inverse = 0; DiffusionDataset.__getitem__, datamodules.py:62) or x * std + mean (inverse = 1). */
int fd_standardise(const float *x_dev, const float *mean_dev, const float *std_dev, float *out_dev, int64_t n, int32_t n_features, int32_t inverse,
                   int32_t device, void *stream);

/* ---- introspection (bench / tests) ---------------------------------------------------------------------- */
/* Number of kernels this library has launched on behalf of `h` since creation (fd_dft/fd_idft count on a global). */
int64_t fd_launch_count(const fd_handle *h);
int64_t fd_global_launch_count(void);
/* Which kernel family fd_score dispatches to for this handle: 0 = generic fp32 kernels, 1 = tcgen05 tensor-core path (transformer,
 * fp16 / TF32 operands, fp32 accumulate), 2 = LSTM sampler kernel on warp-level MMAs (fp16 operands, fp32 accumulate, MUFU tanh gates). */
int fd_active_path(const fd_handle *h);
/* Tuning knobs of a handle (introspection and tests; defaults are the production settings).  Known options:
 *   "attn_bounded_softmax"  1 (default): attention heads whose scores are provably bounded (max|q| * max|k| <= 14 in log2 units,
 *                           checked per series and head at run time) exponentiate without a row maximum; 0: always the exact
 *                           two-pass softmax.  Both give the same result up to rounding.
 *   "persistent_stack"      1 (default): all encoder layers of a score evaluation run as ONE persistent kernel (task queue +
 *                           dependency counters, csrc/fd_step.cu); 0: the per-layer kernels (2 launches per layer) — the cross-check
 *                           path, and the one to use under tools that serialise or replay individual kernels per layer.
 *   "stack_lag"             queue order of the persistent kernel: FFN tasks trail the attention tasks by this many series
 *                           (-1 = batch / 2, the default).
 *   "stack_lanes"           persistent kernel, fd_sample only: the batch is cut into this many sub-batches whose stack kernels are in flight on
 *                           separate streams, each with its own task queue and dependency counters (1..4; 0 = by batch size, the default:
 *                           lanes of at least 340 series, at most 3).  Samples are bit-identical whatever the value.
 *   "stack_debug"           1: per-CTA cycle counters in the persistent kernel (fd_debug_stack_stats); default 0.
 *   "lanes"                 per-layer kernels only: independent sub-batches in flight on separate streams (1..4, default 2).
 *   "fuse_boundary"         1 (default): unembed + scheduler step + embed of the next step in one kernel (for 12 channels and d_model 72
 *                           the variant with the weights as constant operands); 2: the same with the weights in shared memory (any
 *                           shape); 0: three kernels.  All three give bit-identical samples.
 *   "lstm_persistent"       LSTM score network: 1 (default): fd_sample runs the WHOLE reverse-diffusion loop in one kernel launch
 *                           (csrc/fd_lstm.cu; a CTA keeps its series on chip for all steps); 0: one launch per score evaluation + one per
 *                           scheduler step — bit-identical results, the cross-check path.
 *   "lstm_debug"            timing probes of the LSTM kernel (bit mask, tools/lstm_probe.py); any non-zero value gives WRONG results.
 * Unknown names are an error.  (Environment variables FD_ATTN_BOUNDED, FD_STACK, FD_STACK_LAG, FD_STACK_LANES, FD_LANES, FD_FUSE_BOUNDARY preset the
 * same options when a handle is created — a bring-up convenience.) */
int fd_set_option(fd_handle *h, const char *name, int32_t value);
/* Host-only helper (no device needed): the task queue the persistent encoder-stack kernel walks for `batch` series of length
 * `max_len` — entry = bit 31: FFN task (else attention) | bits 24..30: layer | bits 0..23: tile index (FFN) or series * 4 + head group.
 * Writes at most `cap` entries to `out` (may be NULL) and returns the queue length (< 0: bad argument).  Every task's dependencies
 * precede it in the queue; tests/test_host_logic.py checks exactly that. */
/* Bring-up aid: with option "stack_debug" = 1 the persistent kernel accumulates per-CTA cycle counters; this copies them out
 * (64 int64 per CTA: [0..8) lifetime, ATT tasks, ATT cycles, FFN tasks, FFN cycles, ATT / FFN dependency-wait cycles, SM id;
 * [8..36) sums of the ATT task's phase timestamps, [36..56) of the FFN task's — see csrc/fd_step.cu) and clears them.  Returns the number of CTAs written.  Synchronises the device. */
int fd_debug_stack_stats(fd_handle *h, int64_t *out, int32_t cap_ctas);
/* Post-mortem of a protocol time-out inside the persistent kernel.  Every wait in it is bounded: instead of hanging the GPU a
 * stuck wait writes a record to pinned host memory and traps, which surfaces as a CUDA "launch failure" and — like any device trap —
 * invalidates the process's CUDA context.  out8: [0] 1 = mbarrier wait / 2 = dependency-counter wait, [1] CTA, [2] thread,
 * [3..6] barrier address + parity, or task tag + counter value + target.  Returns 1 if a record exists. */
int fd_debug_abort_record(int32_t *out8);
int fd_stack_task_table(int32_t batch, int32_t max_len, int32_t num_layers, int32_t lag, uint32_t *out, int32_t cap);
/* Enable per-kernel CUDA-event timing of the next fd_sample call (adds events around each kernel family); read the
 * accumulated milliseconds afterwards with fd_profile_ms("ffn"|"attn"|"qkv"|"embed"|"unembed_step"|...). */
int fd_profile_enable(fd_handle *h, int32_t enable);
double fd_profile_ms(fd_handle *h, const char *family);
int64_t fd_profile_launches(fd_handle *h, const char *family);

#ifdef __cplusplus
}
#endif
#endif /* FDIFF_B200_H */
