// Shared declarations of the fdiff_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/fdiff_b200.h"

namespace fd {

// ---- error plumbing -------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);

#define FD_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            fd::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

#define FD_CHECK(cond, ...)               \
    do {                                  \
        if (!(cond)) {                    \
            fd::set_error(__VA_ARGS__);   \
            return 1;                     \
        }                                 \
    } while (0)

#define FD_TRY(expr)            \
    do {                        \
        int _r = (expr);        \
        if (_r) return _r;      \
    } while (0)

// ---- per-family profiling (CUDA events on the launching stream) -------------------------------------------------
struct ProfileFamily {
    std::vector<cudaEvent_t> starts, stops;
    int64_t launches = 0;  // launches covered by the recorded event pairs
    double ms = 0.0;       // resolved on demand
};

struct Profiler {
    bool enabled = false;
    std::map<std::string, ProfileFamily> fam;
    void begin(const char *name, cudaStream_t s);
    void end(const char *name, cudaStream_t s, int launches);
    void resolve();
    void clear();
};

// ---- weights --------------------------------------------------------------------------------------------------
struct DevTensor {
    float *ptr = nullptr;
    int64_t numel = 0;
};

struct TransformerLayerW {
    const float *in_w, *in_b, *out_w, *out_b, *l1_w, *l1_b, *l2_w, *l2_b, *n1_w, *n1_b, *n2_w, *n2_b;
    // packed images for the tensor-core path (built by fd_finalize_weights; nullptr on the generic path)
    const float *l1_pack = nullptr, *l2_pack = nullptr, *in_pack = nullptr, *in_bias_pack = nullptr, *out_pack = nullptr,
                *out_pack16 = nullptr,  // out_proj as the fp16 image of the fused FFN-layer kernel
                *in_pack16 = nullptr,   // in_proj per 3-head group as fp16 images [10][80][8 halfs] (persistent encoder-stack kernel)
                *in_pack_half = nullptr, *in_bias_pack_half = nullptr;  // in_proj per 6-head half (fd_attn_stream.cu, max_len > 256)
};
struct LstmLayerW {
    const float *w_ih, *w_hh, *b_ih, *b_hh;
};
struct MlpLayerW {
    const float *w0, *b0, *w3, *b3;
};

}  // namespace fd

#define FD_MAX_LANES 4

// The opaque handle of the C ABI.
struct fd_handle {
    fd_config cfg;
    int finalized = 0;
    int active_path = 0;  // 0 generic fp32, 1 TF32 tensor-core
    std::map<std::string, fd::DevTensor> weights;
    // resolved views
    const float *pos = nullptr, *time_W = nullptr, *time_dw = nullptr, *time_db = nullptr;
    const float *emb_w = nullptr, *emb_b = nullptr, *unemb_w = nullptr, *unemb_b = nullptr;
    std::vector<fd::TransformerLayerW> tl;
    std::vector<fd::LstmLayerW> ll;
    std::vector<fd::MlpLayerW> ml;
    std::vector<float *> owned;  // extra device allocations (packed weights, G, ...)
    float *G = nullptr;           // (L,) diffusion scaling, sde.py:42-60
    // workspace, sized for cap_batch series
    int cap_batch = 0;
    float *ws_x = nullptr, *ws_h = nullptr, *ws_h2 = nullptr, *ws_qkv = nullptr, *ws_att = nullptr, *ws_hid = nullptr,
          *ws_score = nullptr;
    float *ws_himg = nullptr;   // tensor-core path: per series the token rows as the attention kernel's tf32 operand image [18][256][4]
    float *ws_attimg = nullptr; // tensor-core path: per 256-token tile the attention output as the FFN kernel's fp16 operand image [9][256][8]
    float *ws_temb = nullptr;   // (cap_steps, D) time-embedding rows, one per diffusion step
    int temb_per_series = 0;    // set by fd_score_t for the duration of the call: `temb_row` of the score paths is a (batch, D) table, one row per series
    float *ws_tsteps = nullptr; // (cap_steps,) fp32 timesteps on the device
    float *ws_coef = nullptr;   // (cap_steps, 2) fp32 {drift coefficient on x, diffusion scalar} per step
    int cap_steps = 0;
    int attn_fast = 0;          // 1: QKV / attention / out-proj run on the tensor-core kernels too
    int attn_stream = 0;        // 1: max_len > 256 — projection-to-images + streaming attention kernels (fd_attn_stream.cu)
    float *ws_qimg = nullptr, *ws_kvimg = nullptr;  // streaming attention: q and k|v operand images of the batch
    unsigned *ws_nrm = nullptr;                      // streaming attention: per (series, head) max |q|^2, max |k|^2 (float bits)
    int attn_bounded = 1;       // fd_set_option("attn_bounded_softmax"): bounded heads skip the row maximum (env FD_ATTN_BOUNDED=0 turns it off globally)
    int himg_primed = 0;        // 1: ws_himg holds the embedded rows of the step about to run (written by the step-boundary kernel)
    void *bw_host = nullptr;    // step-boundary kernel, constant-operand variant: host copy of the unembed / embed weights (passed by value)
    int bw_ready = 0;
    int himg_fp16 = 0;          // format of that image: 1 = fp16 [10][256][8 halfs] (encoder-stack kernel), 0 = tf32 [18][256][4] (per-layer kernels)
    cudaStream_t lane_stream[FD_MAX_LANES] = {};  // fd_sample: independent sub-batches in flight on separate streams (fills partial waves)
    cudaEvent_t lane_event[FD_MAX_LANES + 1] = {};
    float *stage_noise = nullptr;  // device staging for fd_sample_host
    size_t stage_noise_bytes = 0;
    float *stage_out = nullptr;
    size_t stage_out_bytes = 0;
    // persistent encoder-stack kernel (fd_step.cu): all layers of a score evaluation in one launch
    int stack_enabled = 1;      // fd_set_option("persistent_stack"): 0 = the per-layer kernels (2 launches per layer)
    int stack_lag = -1;         // fd_set_option("stack_lag"): FFN tasks trail the ATT tasks by this many series in the queue (-1: batch / 2)
    int stk_grid = 0;           // resident CTAs (occupancy x SMs)
    uint32_t *stk_table = nullptr;   // task queue for stk_table_batch series
    unsigned *stk_counters = nullptr;  // [0] claim counter | [32 ..) per-tile ATT completions | per-series FFN completions
    int stk_n_tasks = 0, stk_table_batch = 0, stk_table_lag = -2;
    unsigned stk_claims = 0;    // claims consumed by earlier launches
    unsigned stk_k = 0;         // encoder layers completed by earlier launches since the counters were zeroed
    // The queue / counter state above belongs to ONE sequence of launches.  fd_sample with "stack_lanes" > 1 keeps one state per sub-batch
    // (slot k + 1 for lane k, slot 0 for un-split launches) and swaps the live fields with stack_select_slot before each launch.
    struct StkSlot {
        uint32_t *table = nullptr;
        unsigned *counters = nullptr;
        int n_tasks = 0, table_batch = 0, table_lag = -2;
        unsigned claims = 0, k = 0;
    } stk_slots[FD_MAX_LANES + 1];
    int stk_slot = 0;           // which slot the live fields belong to
    int stack_lanes = 0;        // fd_set_option("stack_lanes"): sub-batches of fd_sample whose stack kernels are in flight on separate streams (0 = by batch size)
    int stack_flags = 0;        // fd_set_option("stack_flags"): bring-up switches of the stack kernel
    int stack_debug = 0;        // fd_set_option("stack_debug"): per-CTA cycle counters of the stack kernel (fd_debug_stack_stats)
    long long *stk_dbg = nullptr;
    int lanes = 2;              // fd_set_option("lanes"): half-batches in flight on separate streams (per-layer kernels only)
    int fuse_boundary = 1;      // fd_set_option("fuse_boundary"): unembed + scheduler step + embed in one kernel
    void *lstm_wfrag = nullptr; // LSTM sampler kernel (fd_lstm.cu): per layer the [W_ih | W_hh] A fragments, fp16, gate rows permuted
    float *lstm_bias = nullptr; //                                   per layer b_ih + b_hh
    int lstm_debug = 0;         // fd_set_option("lstm_debug"): timing probes of the LSTM kernel (wrong results)
    int lstm_persistent = 1;    // fd_set_option("lstm_persistent"): 0 = fd_sample launches one score evaluation + one scheduler step per diffusion step
    int64_t launches = 0;
    fd::Profiler prof;
    int prof_requested = 0;  // 0 = off, n = profile every n-th diffusion step of fd_sample
};

namespace fd {

int ensure_workspace(fd_handle *h, int batch, int n_steps, cudaStream_t s);

// ---- kernels: generic fp32 path (fd_generic.cu) -----------------------------------------------------------------
// Y[M,N] = act( X[M,K] · W[N,K]^T + bias[N] + rowtab[(m % rowtab_period), N] + vec[N] + residual[M,N] )
struct GemmEpilogue {
    const float *bias = nullptr;      // (N)
    const float *rowtab = nullptr;    // (rowtab_period, N): positional table, indexed by m % period
    int rowtab_period = 1;
    const float *vec = nullptr;       // (N): time-embedding row shared by every row, or with vec_rows > 0 a table (M / vec_rows, N)
    int vec_rows = 0;                 //      whose row m / vec_rows is added to output row m (per-series diffusion times, fd_score_t)
    const float *residual = nullptr;  // (M, N)
    int relu = 0;
};
int launch_gemm(fd_handle *h, const float *X, const float *W, float *Y, int M, int N, int K, const GemmEpilogue &ep,
                cudaStream_t s);
int launch_add_layernorm(fd_handle *h, const float *a, const float *w, const float *b, float *y, int M, int D,
                         cudaStream_t s);
int launch_attention(fd_handle *h, const float *qkv, float *out, int B, int L, int D, int H, cudaStream_t s);
int launch_time_embedding(fd_handle *h, const float *tsteps_dev, int n, float *temb, cudaStream_t s);
int launch_time_embedding_scalar(fd_handle *h, float t, float *temb, cudaStream_t s);
int launch_lstm_layer(fd_handle *h, const float *xin, const float *w_hh, const float *b_hh, float *u, int B, int L, int D,
                      cudaStream_t s);
int launch_sde_step(fd_handle *h, const float *x, const float *score, const float *z, float *out, int B, float cx, float d0,
                    float dt, float sqrt_dt, uint64_t seed, uint64_t first_series, uint32_t draw, cudaStream_t s);
int launch_prior(fd_handle *h, const float *z, float *out, int B, uint64_t seed, uint64_t first_series, cudaStream_t s);
int launch_normal(fd_handle *h, float *out, int B, uint64_t seed, uint64_t first_series, uint32_t draw, cudaStream_t s);

// sampler-loop building blocks of the transformer path
int transformer_embed(fd_handle *h, const float *x, const float *temb_row, int B, cudaStream_t s);
int transformer_layers(fd_handle *h, int B, cudaStream_t s);
int step_boundary_supported(const fd_handle *h);
int launch_step_boundary(fd_handle *h, float *hbuf, float *x, const float *z, const float *temb_next, int B, float cx, float d0, float dt,
                         float sqrt_dt, uint64_t seed, uint64_t first_series, uint32_t draw, int do_embed, cudaStream_t s);

// score network drivers
int attention_block(fd_handle *h, int layer, float *hbuf, int B, cudaStream_t s);  // LN1(h + out_proj(MHA(h))) in place, either path
int ffn_block(fd_handle *h, int layer, float *hbuf, int M, cudaStream_t s);  // LN2(h + FFN(h)) in place, either path
int score_generic(fd_handle *h, const float *x, const float *temb_row, float *score, int B, cudaStream_t s);
int fast_path_supported(const fd_config &cfg);
int fast_finalize(fd_handle *h);
int attn_path_supported(const fd_config &cfg);
int attn_finalize(fd_handle *h);
int attn_dump_tlog();
int ffn_dump_tlog();
// himg (nullable): the series' token rows as the tf32 operand image (else gathered from hbuf); att_img (nullable): write the fp16 operand
// image of launch_outproj_ffn_fast instead of fp32 rows to att_out
int launch_attention_fast(fd_handle *h, int layer, const float *hbuf, const float *himg, float *att_out, void *att_img, int B, cudaStream_t s);
int lstm_stack_tc_supported(const fd_handle *h);
int lstm_tc_finalize(fd_handle *h);
int lstm_generic_finalize(fd_handle *h);
// LSTM score network / whole reverse-diffusion loop in one launch (fd_lstm.cu): n_steps scheduler steps on x, or (score_out != nullptr) one
// score evaluation of x
int launch_lstm_sampler(fd_handle *h, float *x, float *score_out, const float *temb, const float *coef, const float *noise, int B, int n_steps,
                        float dt, float sqrt_dt, uint64_t seed, uint64_t first_series, cudaStream_t s);
int attn_stream_supported(const fd_config &cfg);
int attn_stream_finalize(fd_handle *h);
size_t stream_qimg_floats(int B, int L);
size_t stream_kvimg_floats(int B, int L);
size_t stream_nrm_words(int B);
int launch_attention_stream(fd_handle *h, int layer, const float *hbuf, float *att_out, void *att_img, int B, cudaStream_t s);
int launch_outproj_ln_fast(fd_handle *h, int layer, const float *att_in, float *hbuf, int B, cudaStream_t s);
// LN2(FFN(LN1(h + out_proj(att)))); himg_out (nullable): also leave the result as the next attention kernel's tf32 operand image
int launch_outproj_ffn_fast(fd_handle *h, int layer, const void *att_img, float *hbuf, int M, float *himg_out, cudaStream_t s);
int launch_ffn_fast(fd_handle *h, int layer, float *hbuf, int M, cudaStream_t s);  // h <- LN2(h + FFN(h)), tcgen05 TF32
// persistent encoder-stack kernel (fd_step.cu)
int stack_supported(const fd_handle *h);
void stack_select_slot(fd_handle *h, int slot);  // swap the live queue / counter state of the handle (fd_sample lanes)
void stack_select_slot(fd_handle *h, int slot);
int stack_finalize(fd_handle *h);
int launch_encoder_stack(fd_handle *h, int B, cudaStream_t s);  // ws_h <- all encoder layers(ws_h), one launch

// ---- FFT (fd_fft.cu) -------------------------------------------------------------------------------------------
int launch_dft(const float *x, float *out, int B, int L, int C, const float *mean, const float *std, bool inverse,
               cudaStream_t s);

int launch_spectral_density(const float *packed, float *out, int B, int L, int C, cudaStream_t s);

// ---- Wasserstein metrics (fd_wass.cu) ---------------------------------------------------------------------------
int launch_wasserstein(const float *x, const float *y, const double *dirs, int n, int m, int d, int K, int standardise, float *work, double *out,
                       cudaStream_t s);
size_t wasserstein_work_bytes(int n, int m, int K);
int launch_feature_stats(const float *x, float *mean, float *stdv, long long n, int F, double *work, cudaStream_t s);
size_t feature_stats_work_bytes(int F);
int launch_standardise(const float *x, const float *mean, const float *stdv, float *out, long long n, int F, int inverse, cudaStream_t s);

extern int64_t g_global_launches;

}  // namespace fd
