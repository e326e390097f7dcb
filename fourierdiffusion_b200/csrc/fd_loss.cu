// Denoising score-matching loss, forward only — the evaluation half of get_sde_loss_fn (src/fdiff/utils/losses.py:39-125), i.e. what
// ScoreModule.validation_step computes (score_models.py:110-113): perturb the batch with the SDE's transition kernel at PER-SERIES
// diffusion times, evaluate the score network on it, and reduce the weighted squared error.  SURVEY.md §8(f) rank 4, forward part;
// the backward pass / optimiser are out of scope.
//
//   marginal_kernel    per series b: mean coefficient m_b and std scalar s_b of p(x_t | x_0) at t_b      (sde.py:108-123 VE, :187-210 VP)
//   perturb_kernel     x_t[b,l,c] = m_b x_0[b,l,c] + (s_b G_l) z[b,l,c]                                  (losses.py:63-84, sde.py:66-77)
//   (score network)    fd_score_t: the same kernels as fd_score with one time-embedding row per series   (score_models.py:67-94)
//   loss_series_kernel loss_b = w_b reduce_{l,c} (score + z / std)^2,  w_b = 1 / sum_l 1 / std_{b,l}^2   (losses.py:89-104)
//                      or reduce (std (score + z / std))^2 with likelihood weighting                     (losses.py:106-118)
//   loss_mean_kernel   loss = mean_b loss_b (fp64 sum, one CTA: deterministic)                           (losses.py:120)
//
// Everything but the score network is one pass over (batch, L, C) fp32: HBM-bound and tiny next to the network.
#include "fd_common.cuh"

namespace fd {

#define FDL_LAUNCH_CHECK()                                                                                 \
    do {                                                                                                   \
        cudaError_t _e = cudaGetLastError();                                                               \
        if (_e != cudaSuccess) {                                                                           \
            fd::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return 1;                                                                                      \
        }                                                                                                  \
    } while (0)

static inline void loss_count_launch(fd_handle *h, int n = 1) {
    h->launches += n;
    g_global_launches += n;
}

// coef[2 b] = mean coefficient, coef[2 b + 1] = std scalar (std_{b,l} = scalar * G_l), fp32 in the reference's operation order
__global__ void __launch_bounds__(256) marginal_kernel(const float *__restrict__ t, float *__restrict__ coef, int B, int is_ve, float p0, float p1,
                                                       float p_ratio_or_diff) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float tb = t[b];
    float m, s;
    if (is_ve) {  // sde.py:117-122: std = sigma_min * (sigma_max / sigma_min) ** t, mean = x
        m = 1.0f;
        s = __fmul_rn(p0, powf(p_ratio_or_diff, tb));
    } else {  // sde.py:196-208: log_mean_coeff = -0.25 t^2 (beta_1 - beta_0) - 0.5 t beta_0
        const float a = __fmul_rn(__fmul_rn(-0.25f, __fmul_rn(tb, tb)), p_ratio_or_diff);
        const float c = __fmul_rn(__fmul_rn(0.5f, tb), p0);
        const float lmc = __fsub_rn(a, c);
        m = expf(lmc);
        s = sqrtf(__fsub_rn(1.0f, expf(__fmul_rn(2.0f, lmc))));
    }
    coef[2 * b] = m;
    coef[2 * b + 1] = s;
}

__global__ void __launch_bounds__(256) perturb_kernel(const float *__restrict__ x0, const float *__restrict__ z, const float *__restrict__ coef,
                                                      const float *__restrict__ G, float *__restrict__ out, long long n, int L, int C) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int LC = L * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long b = i / LC;
        const int l = (int)(i - b * LC) / C;
        const float m = coef[2 * b], std = __fmul_rn(coef[2 * b + 1], G[l]);
        out[i] = __fadd_rn(__fmul_rn(m, x0[i]), __fmul_rn(std, z[i]));  // mean + diag(std) z, no contraction (losses.py:73,82-84)
    }
}

__device__ __forceinline__ double block_sum(double v, double *sm) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[w] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < nw; ++i) r += sm[i];
    return r;
}

// one CTA per series
__global__ void __launch_bounds__(256) loss_series_kernel(const float *__restrict__ score, const float *__restrict__ z, const float *__restrict__ coef,
                                                          const float *__restrict__ G, float *__restrict__ losses, int L, int C, int likelihood,
                                                          int reduce_mean) {
    __shared__ double sm[8];
    const int b = blockIdx.x, LC = L * C;
    const float s = coef[2 * b + 1];
    float w = 1.0f;
    if (!likelihood) {  // losses.py:93: 1 / tr(Sigma^-1)
        double inv = 0.0;
        for (int l = threadIdx.x; l < L; l += blockDim.x) {
            const float std = __fmul_rn(s, G[l]);
            inv += (double)(1.0f / __fmul_rn(std, std));
        }
        w = 1.0f / (float)block_sum(inv, sm);
    }
    const float *sc = score + (size_t)b * LC, *zb = z + (size_t)b * LC;
    double acc = 0.0;
    for (int i = threadIdx.x; i < LC; i += blockDim.x) {
        const float std = __fmul_rn(s, G[i / C]);
        const float diff = __fadd_rn(sc[i], __fmul_rn(1.0f / std, zb[i]));  // score + Sigma^{-1/2} z
        float e;
        if (likelihood) {
            const float sd = __fmul_rn(std, diff);
            e = __fmul_rn(sd, sd);
        } else {
            e = __fmul_rn(w, __fmul_rn(diff, diff));
        }
        acc += (double)e;
    }
    const double tot = block_sum(acc, sm);
    if (threadIdx.x == 0) losses[b] = (float)(reduce_mean ? tot / (double)LC : 0.5 * tot);  // losses.py:34-38
}

__global__ void __launch_bounds__(256) loss_mean_kernel(const float *__restrict__ losses, float *__restrict__ loss, int B) {
    __shared__ double sm[8];
    double acc = 0.0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) acc += (double)losses[i];
    const double tot = block_sum(acc, sm);
    if (threadIdx.x == 0) *loss = (float)(tot / (double)B);
}

static int launch_marginal(fd_handle *h, const float *t_dev, float *coef, int B, cudaStream_t s) {
    const fd_config &c = h->cfg;
    const int ve = c.sched_kind == FD_SCHED_VE;
    // the reference forms sigma_max / sigma_min on fp32 tensors (sde.py:117-119) and beta_1 - beta_0 on Python floats (sde.py:197)
    const float p0 = (float)c.sched_p0, p1 = (float)c.sched_p1;
    const float third = ve ? p1 / p0 : (float)(c.sched_p1 - c.sched_p0);
    marginal_kernel<<<(B + 255) / 256, 256, 0, s>>>(t_dev, coef, B, ve, p0, p1, third);
    FDL_LAUNCH_CHECK();
    loss_count_launch(h);
    return 0;
}

static int launch_perturb(fd_handle *h, const float *x0, const float *z, const float *coef, float *out, int B, cudaStream_t s) {
    const fd_config &c = h->cfg;
    const long long n = (long long)B * c.max_len * c.n_channels;
    const long long blocks = (n + 255) / 256;
    perturb_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, s>>>(x0, z, coef, h->G, out, n, c.max_len, c.n_channels);
    FDL_LAUNCH_CHECK();
    loss_count_launch(h);
    return 0;
}

// the (batch, 2) coefficient table lives in the time-step workspace (ws_coef holds 2 floats per "step"; sized by ensure_workspace(h, batch, batch))
static int score_per_series(fd_handle *h, const float *x_dev, const float *t_dev, float *score_dev, int batch, cudaStream_t s) {
    FD_TRY(launch_time_embedding(h, t_dev, batch, h->ws_temb, s));  // (batch, D): GaussianFourierProjection of every series' own time
    h->temb_per_series = 1;
    const int rc = score_generic(h, x_dev, h->ws_temb, score_dev, batch, s);
    h->temb_per_series = 0;
    return rc;
}

}  // namespace fd

using namespace fd;

extern "C" {

int fd_score_t(fd_handle *h, const float *x_dev, const float *t_dev, float *score_dev, int32_t batch, void *stream) {
    FD_CHECK(h && x_dev && t_dev && score_dev && batch > 0, "fd_score_t: bad argument");
    FD_CHECK(h->finalized, "fd_score_t: call fd_finalize_weights first");
    FD_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    FD_TRY(ensure_workspace(h, batch, batch, s));
    return score_per_series(h, x_dev, t_dev, score_dev, batch, s);
}

int fd_perturb(fd_handle *h, const float *x0_dev, const float *t_dev, const float *z_dev, float *out_dev, float *std_scalar_dev, int32_t batch,
               void *stream) {
    FD_CHECK(h && x0_dev && t_dev && z_dev && out_dev && batch > 0, "fd_perturb: bad argument");
    FD_CHECK(h->G, "fd_perturb: the scheduler's G (noise_scheduler.G) has not been set");
    FD_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    FD_TRY(ensure_workspace(h, batch, batch, s));
    FD_TRY(launch_marginal(h, t_dev, h->ws_coef, batch, s));
    FD_TRY(launch_perturb(h, x0_dev, z_dev, h->ws_coef, out_dev, batch, s));
    if (std_scalar_dev)  // column 1 of the (batch, 2) table
        FD_CUDA(cudaMemcpy2DAsync(std_scalar_dev, sizeof(float), h->ws_coef + 1, 2 * sizeof(float), sizeof(float), batch, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int fd_sde_loss(fd_handle *h, const float *x0_dev, const float *t_dev, const float *z_dev, int32_t likelihood_weighting, int32_t reduce_mean,
                float *losses_dev, float *loss_dev, int32_t batch, void *stream) {
    FD_CHECK(h && x0_dev && t_dev && z_dev && loss_dev && batch > 0, "fd_sde_loss: bad argument");
    FD_CHECK(h->finalized, "fd_sde_loss: call fd_finalize_weights first");
    FD_CHECK(h->G, "fd_sde_loss: the scheduler's G (noise_scheduler.G) has not been set");
    FD_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    const fd_config &c = h->cfg;
    FD_TRY(ensure_workspace(h, batch, batch, s));
    FD_TRY(launch_marginal(h, t_dev, h->ws_coef, batch, s));
    FD_TRY(launch_perturb(h, x0_dev, z_dev, h->ws_coef, h->ws_x, batch, s));
    FD_TRY(score_per_series(h, h->ws_x, t_dev, h->ws_score, batch, s));
    // per-series losses: caller's buffer, or the (dead after the embed) time-step array of the workspace (batch + 1 floats)
    float *losses = losses_dev ? losses_dev : h->ws_tsteps;
    loss_series_kernel<<<batch, 256, 0, s>>>(h->ws_score, z_dev, h->ws_coef, h->G, losses, c.max_len, c.n_channels, likelihood_weighting != 0,
                                             reduce_mean != 0);
    FDL_LAUNCH_CHECK();
    loss_mean_kernel<<<1, 256, 0, s>>>(losses, loss_dev, batch);
    FDL_LAUNCH_CHECK();
    loss_count_launch(h, 2);
    return 0;
}

}  // extern "C"
