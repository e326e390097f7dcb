// C-ABI entry points of libfdiff_b200 (see include/fdiff_b200.h for the contract and the reference call sites).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "fd_common.cuh"

namespace fd {

static thread_local char g_err[1024] = "";
int64_t g_global_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- profiler -----------------------------------------------------------------------------------------------------
void Profiler::begin(const char *name, cudaStream_t s) {
    if (!enabled) return;
    ProfileFamily &f = fam[name];
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f.starts.push_back(a);
    f.stops.push_back(b);
    cudaEventRecord(a, s);
}
void Profiler::end(const char *name, cudaStream_t s, int launches) {
    if (!enabled) return;
    ProfileFamily &f = fam[name];
    cudaEventRecord(f.stops.back(), s);
    f.launches += launches;
}
void Profiler::resolve() {
    for (auto &kv : fam) {
        ProfileFamily &f = kv.second;
        for (size_t i = 0; i < f.starts.size(); ++i) {
            cudaEventSynchronize(f.stops[i]);
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, f.starts[i], f.stops[i]) == cudaSuccess) f.ms += ms;
            cudaEventDestroy(f.starts[i]);
            cudaEventDestroy(f.stops[i]);
        }
        f.starts.clear();
        f.stops.clear();
    }
}
void Profiler::clear() {
    resolve();
    fam.clear();
}

// ---- workspace ----------------------------------------------------------------------------------------------------
static int dev_alloc(float **p, size_t n_floats) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    if (n_floats == 0) return 0;
    FD_CUDA(cudaMalloc((void **)p, n_floats * sizeof(float)));
    return 0;
}

int ensure_workspace(fd_handle *h, int batch, int n_steps, cudaStream_t s) {
    const fd_config &c = h->cfg;
    const size_t L = c.max_len, C = c.n_channels, D = c.d_model;
    if (batch > h->cap_batch) {
        h->cap_batch = 0;  // a failed growth must not leave a stale capacity next to freed / half-replaced buffers
        size_t M = (size_t)batch * L;
        size_t wide = 3 * D, hid = 0;
        if (c.model_kind == FD_MODEL_LSTM) wide = 4 * D;
        if (c.model_kind == FD_MODEL_TRANSFORMER) hid = M * (size_t)c.d_ff;
        if (c.model_kind == FD_MODEL_MLP) hid = (size_t)batch * (size_t)c.d_ff;
        FD_TRY(dev_alloc(&h->ws_x, batch * L * C));
        FD_TRY(dev_alloc(&h->ws_score, batch * L * C));
        FD_TRY(dev_alloc(&h->ws_h, (M + 128) * D));  // + one 128-token tile: the stack kernel stages whole tiles of rows
        FD_TRY(dev_alloc(&h->ws_h2, M * D));
        FD_TRY(dev_alloc(&h->ws_att, M * D));
        FD_TRY(dev_alloc(&h->ws_qkv, M * wide));
        if (h->active_path == 0 || c.model_kind != FD_MODEL_TRANSFORMER) FD_TRY(dev_alloc(&h->ws_hid, hid));
        if (h->active_path == 1 && h->attn_fast) {
            // operand images exchanged by the two layer kernels; zero-filled once: positions >= max_len of a series image and the rows
            // past the last token of the last tile are never written
            const size_t himg = h->attn_stream ? 0 : (size_t)batch * 18 * 256 * 4, tiles = (M + 255) / 256 + FD_MAX_LANES + 1,
                         attimg = tiles * 9 * 256 * 4;
            FD_TRY(dev_alloc(&h->ws_himg, himg));
            FD_TRY(dev_alloc(&h->ws_attimg, attimg));
            // ordered on the caller's stream (the legacy default stream does not synchronise with non-blocking streams)
            if (himg) FD_CUDA(cudaMemsetAsync(h->ws_himg, 0, himg * sizeof(float), s));
            FD_CUDA(cudaMemsetAsync(h->ws_attimg, 0, attimg * sizeof(float), s));
            if (h->attn_stream) {  // q / k|v operand images of the whole batch (every padded position is rewritten by each projection launch)
                FD_TRY(dev_alloc(&h->ws_qimg, stream_qimg_floats(batch, c.max_len)));
                FD_TRY(dev_alloc(&h->ws_kvimg, stream_kvimg_floats(batch, c.max_len)));
                float *nrm = (float *)h->ws_nrm;
                FD_TRY(dev_alloc(&nrm, stream_nrm_words(batch)));
                h->ws_nrm = (unsigned *)nrm;
            }
        }
        h->cap_batch = batch;
    }
    if (n_steps > h->cap_steps) {
        FD_TRY(dev_alloc(&h->ws_temb, (size_t)(n_steps + 1) * D));
        FD_TRY(dev_alloc(&h->ws_tsteps, (size_t)n_steps + 1));
        FD_TRY(dev_alloc(&h->ws_coef, (size_t)2 * n_steps));
        h->cap_steps = n_steps;
    }
    return 0;
}

int launch_time_embedding_scalar(fd_handle *h, float t, float *temb, cudaStream_t s);

static const float *find_w(fd_handle *h, const std::string &name, int64_t numel, bool at_least = false) {
    auto it = h->weights.find(name);
    if (it == h->weights.end()) {
        set_error("missing weight '%s'", name.c_str());
        return nullptr;
    }
    if (at_least ? it->second.numel < numel : it->second.numel != numel) {
        set_error("weight '%s' has %lld elements, expected %s%lld", name.c_str(), (long long)it->second.numel,
                  at_least ? ">= " : "", (long long)numel);
        return nullptr;
    }
    return it->second.ptr;
}

}  // namespace fd

using namespace fd;

extern "C" {

int fd_abi_version(void) { return FD_ABI_VERSION; }
const char *fd_last_error(void) { return fd::g_err; }

int fd_create(const fd_config *cfg, fd_handle **out) {
    FD_CHECK(cfg && out, "fd_create: null argument");
    FD_CHECK(cfg->struct_size == (int32_t)sizeof(fd_config), "fd_create: fd_config size mismatch (%d vs %d)", cfg->struct_size,
             (int)sizeof(fd_config));
    FD_CHECK(cfg->max_len > 0 && cfg->n_channels > 0 && cfg->d_model > 0 && cfg->num_layers >= 0, "fd_create: bad shape");
    FD_CHECK(cfg->model_kind >= FD_MODEL_TRANSFORMER && cfg->model_kind <= FD_MODEL_MLP, "fd_create: unknown model kind %d",
             cfg->model_kind);
    FD_CHECK(cfg->sched_kind == FD_SCHED_VP || cfg->sched_kind == FD_SCHED_VE, "fd_create: Scheduler not recognized (%d)",
             cfg->sched_kind);
    if (cfg->model_kind == FD_MODEL_TRANSFORMER)
        FD_CHECK(cfg->n_head > 0 && cfg->d_model % cfg->n_head == 0 && cfg->d_ff > 0, "fd_create: d_model %d not divisible by n_head %d",
                 cfg->d_model, cfg->n_head);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    FD_CHECK(e == cudaSuccess && ndev > 0, "fd_create: no CUDA device available (%s) — this library has no CPU fallback",
             cudaGetErrorString(e));
    FD_CHECK(cfg->device >= 0 && cfg->device < ndev, "fd_create: device %d out of range (%d devices)", cfg->device, ndev);
    FD_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    FD_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    FD_CHECK(prop.major == 10, "fd_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major,
             prop.minor);
    fd_handle *h = new fd_handle();
    h->cfg = *cfg;
    // environment overrides of the handle options (bring-up convenience; fd_set_option is the API)
    if (getenv("FD_ATTN_BOUNDED")) h->attn_bounded = atoi(getenv("FD_ATTN_BOUNDED")) != 0;
    if (getenv("FD_STACK")) h->stack_enabled = atoi(getenv("FD_STACK")) != 0;
    if (getenv("FD_STACK_LAG")) h->stack_lag = atoi(getenv("FD_STACK_LAG"));
    if (getenv("FD_STACK_FLAGS")) h->stack_flags = atoi(getenv("FD_STACK_FLAGS"));
    if (getenv("FD_STACK_LANES")) h->stack_lanes = std::max(0, std::min(FD_MAX_LANES, atoi(getenv("FD_STACK_LANES"))));
    if (getenv("FD_LANES")) h->lanes = std::max(1, std::min(FD_MAX_LANES, atoi(getenv("FD_LANES"))));
    if (getenv("FD_FUSE_BOUNDARY")) h->fuse_boundary = atoi(getenv("FD_FUSE_BOUNDARY"));
    // default G (sde.py:42-60) in fp32; the host mirror overrides it with the tensor its scheduler holds ("noise_scheduler.G")
    std::vector<float> G(cfg->max_len, 1.0f);
    if (cfg->fourier_noise_scaling) {
        float inv = (float)(1.0 / sqrt(2.0));
        float s2 = (float)sqrt(2.0);
        for (auto &g : G) g = inv * 1.0f;
        G[0] = G[0] * s2;
        if (cfg->max_len % 2 == 0) G[cfg->max_len / 2] = G[cfg->max_len / 2] * s2;
    }
    if (cudaMalloc((void **)&h->G, G.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(h->G, G.data(), G.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("fd_create: allocating G failed: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return 1;
    }
    *out = h;
    return 0;
}

int fd_destroy(fd_handle *h) {
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    attn_dump_tlog();
    ffn_dump_tlog();
    h->prof.clear();
    for (auto &kv : h->weights) cudaFree(kv.second.ptr);
    for (float *p : h->owned) cudaFree(p);
    float *bufs[] = {h->G,      h->ws_x,   h->ws_h,      h->ws_h2,   h->ws_qkv,      h->ws_att,   h->ws_hid,
                     h->ws_score, h->ws_temb, h->ws_tsteps, h->ws_coef, h->stage_noise, h->stage_out, h->ws_himg, h->ws_attimg, h->ws_qimg, h->ws_kvimg, (float *)h->ws_nrm};
    for (float *p : bufs)
        if (p) cudaFree(p);
    if (h->bw_host) free(h->bw_host);
    stack_select_slot(h, 0);  // the live state goes back to slot 0; the other slots hold what the lanes allocated
    if (h->stk_table) cudaFree(h->stk_table);
    if (h->stk_counters) cudaFree(h->stk_counters);
    for (int k = 1; k <= FD_MAX_LANES; ++k) {
        if (h->stk_slots[k].table) cudaFree(h->stk_slots[k].table);
        if (h->stk_slots[k].counters) cudaFree(h->stk_slots[k].counters);
    }
    if (h->stk_dbg) cudaFree(h->stk_dbg);
    for (int k = 0; k < FD_MAX_LANES; ++k)
        if (h->lane_stream[k]) cudaStreamDestroy(h->lane_stream[k]);
    for (int k = 0; k <= FD_MAX_LANES; ++k)
        if (h->lane_event[k]) cudaEventDestroy(h->lane_event[k]);
    delete h;
    return 0;
}

int fd_set_weight(fd_handle *h, const char *name, const float *data_host, int64_t numel) {
    FD_CHECK(h && name && data_host && numel > 0, "fd_set_weight: bad argument");
    FD_CUDA(cudaSetDevice(h->cfg.device));
    std::string key(name);
    if (key == "noise_scheduler.G") {
        FD_CHECK(numel == h->cfg.max_len, "fd_set_weight: G has %lld elements, expected max_len=%d", (long long)numel, h->cfg.max_len);
        FD_CUDA(cudaMemcpy(h->G, data_host, numel * sizeof(float), cudaMemcpyHostToDevice));
        return 0;
    }
    DevTensor &t = h->weights[key];
    if (t.ptr && t.numel != numel) {
        cudaFree(t.ptr);
        t.ptr = nullptr;
    }
    if (!t.ptr) FD_CUDA(cudaMalloc((void **)&t.ptr, numel * sizeof(float)));
    t.numel = numel;
    FD_CUDA(cudaMemcpy(t.ptr, data_host, numel * sizeof(float), cudaMemcpyHostToDevice));
    h->finalized = 0;
    h->bw_ready = 0;
    return 0;
}

int fd_finalize_weights(fd_handle *h) {
    FD_CHECK(h, "fd_finalize_weights: null handle");
    FD_CUDA(cudaSetDevice(h->cfg.device));
    const fd_config &c = h->cfg;
    const int64_t D = c.d_model, C = c.n_channels, L = c.max_len, F = c.d_ff;
    const int64_t half = (D + 1) / 2;
#define W_(dst, name, n)                        \
    do {                                        \
        (dst) = find_w(h, (name), (n));         \
        if (!(dst)) return 1;                   \
    } while (0)
    W_(h->time_W, "time_encoder.W", half);
    W_(h->time_dw, "time_encoder.dense.weight", D * D);
    W_(h->time_db, "time_encoder.dense.bias", D);
    h->tl.clear();
    h->ll.clear();
    h->ml.clear();
    if (c.model_kind == FD_MODEL_MLP) {
        W_(h->emb_w, "embedder.weight", D * L * C);
        W_(h->emb_b, "embedder.bias", D);
        W_(h->unemb_w, "unembedder.weight", L * C * D);
        W_(h->unemb_b, "unembedder.bias", L * C);
    } else {
        W_(h->emb_w, "embedder.weight", D * C);
        W_(h->emb_b, "embedder.bias", D);
        W_(h->unemb_w, "unembedder.weight", C * D);
        W_(h->unemb_b, "unembedder.bias", C);
    }
    char nm[256];
    for (int i = 0; i < c.num_layers; ++i) {
        if (c.model_kind == FD_MODEL_TRANSFORMER) {
            TransformerLayerW w;
            auto P = [&](const char *suffix) {
                snprintf(nm, sizeof(nm), "backbone.layers.%d.%s", i, suffix);
                return std::string(nm);
            };
            W_(w.in_w, P("self_attn.in_proj_weight"), 3 * D * D);
            W_(w.in_b, P("self_attn.in_proj_bias"), 3 * D);
            W_(w.out_w, P("self_attn.out_proj.weight"), D * D);
            W_(w.out_b, P("self_attn.out_proj.bias"), D);
            W_(w.l1_w, P("linear1.weight"), F * D);
            W_(w.l1_b, P("linear1.bias"), F);
            W_(w.l2_w, P("linear2.weight"), D * F);
            W_(w.l2_b, P("linear2.bias"), D);
            W_(w.n1_w, P("norm1.weight"), D);
            W_(w.n1_b, P("norm1.bias"), D);
            W_(w.n2_w, P("norm2.weight"), D);
            W_(w.n2_b, P("norm2.bias"), D);
            h->tl.push_back(w);
        } else if (c.model_kind == FD_MODEL_LSTM) {
            LstmLayerW w;
            auto P = [&](const char *suffix) {
                snprintf(nm, sizeof(nm), "backbone.%d.%s", i, suffix);
                return std::string(nm);
            };
            W_(w.w_ih, P("weight_ih_l0"), 4 * D * D);
            W_(w.w_hh, P("weight_hh_l0"), 4 * D * D);
            W_(w.b_ih, P("bias_ih_l0"), 4 * D);
            W_(w.b_hh, P("bias_hh_l0"), 4 * D);
            h->ll.push_back(w);
        } else {
            MlpLayerW w;
            auto P = [&](const char *suffix) {
                snprintf(nm, sizeof(nm), "backbone.%d.%s", i, suffix);
                return std::string(nm);
            };
            W_(w.w0, P("0.weight"), F * D);
            W_(w.b0, P("0.bias"), F);
            W_(w.w3, P("3.weight"), D * F);
            W_(w.b3, P("3.bias"), D);
            h->ml.push_back(w);
        }
    }
    if (c.model_kind == FD_MODEL_TRANSFORMER) {
        h->pos = find_w(h, "pos_encoder.embedding.weight", L * D, /*at_least=*/true);
        if (!h->pos) return 1;
    }
#undef W_
    h->active_path = 0;
    if (c.model_kind == FD_MODEL_LSTM) {
        FD_TRY(lstm_generic_finalize(h));
        FD_TRY(lstm_tc_finalize(h));
    }
    if (c.math_mode == FD_MATH_TF32 && fast_path_supported(c)) {
        FD_TRY(fast_finalize(h));
        h->active_path = 1;
        h->attn_fast = 0;
        h->attn_stream = 0;
        if ((attn_path_supported(c) || attn_stream_supported(c)) && !(getenv("FD_FAST_ATTN") && atoi(getenv("FD_FAST_ATTN")) == 0)) {
            FD_TRY(attn_finalize(h));
            h->attn_fast = 1;
            if (attn_stream_supported(c)) {
                FD_TRY(attn_stream_finalize(h));
                h->attn_stream = 1;
            }
            if (!h->attn_stream) FD_TRY(stack_finalize(h));
        }
    }
    h->finalized = 1;
    h->bw_ready = 0;
    return 0;
}

int fd_active_path(const fd_handle *h) {
    if (!h) return -1;
    if (h->cfg.model_kind == FD_MODEL_LSTM && lstm_stack_tc_supported(h)) return 2;  // fp16 warp-MMA LSTM stack (fd_lstm.cu)
    return h->active_path;
}
int64_t fd_launch_count(const fd_handle *h) { return h ? h->launches : 0; }
int64_t fd_global_launch_count(void) { return fd::g_global_launches; }

int fd_set_option(fd_handle *h, const char *name, int32_t value) {
    FD_CHECK(h && name, "fd_set_option: null argument");
    if (strcmp(name, "attn_bounded_softmax") == 0) {
        h->attn_bounded = value != 0;
        return 0;
    }
    if (strcmp(name, "persistent_stack") == 0) {
        h->stack_enabled = value != 0;
        return 0;
    }
    if (strcmp(name, "stack_flags") == 0) {
        h->stack_flags = value;
        return 0;
    }
    if (strcmp(name, "stack_debug") == 0) {
        h->stack_debug = value != 0;
        return 0;
    }
    if (strcmp(name, "stack_lag") == 0) {
        h->stack_lag = value;
        return 0;
    }
    if (strcmp(name, "stack_lanes") == 0) {
        FD_CHECK(value >= 0 && value <= FD_MAX_LANES, "fd_set_option: stack_lanes must be in 0..%d", FD_MAX_LANES);
        h->stack_lanes = value;
        return 0;
    }
    if (strcmp(name, "lanes") == 0) {
        FD_CHECK(value >= 1 && value <= FD_MAX_LANES, "fd_set_option: lanes must be in [1, %d]", FD_MAX_LANES);
        h->lanes = value;
        return 0;
    }
    if (strcmp(name, "fuse_boundary") == 0) {
        h->fuse_boundary = value;  // 0: three kernels; 1: fused, weights as constant operands where a specialisation exists; 2: fused, weights in shared memory
        return 0;
    }
    if (strcmp(name, "lstm_debug") == 0) {
        h->lstm_debug = value;
        return 0;
    }
    if (strcmp(name, "lstm_persistent") == 0) {
        h->lstm_persistent = value != 0;
        return 0;
    }
    set_error("fd_set_option: unknown option '%s'", name);
    return 1;
}

int fd_profile_enable(fd_handle *h, int32_t enable) {
    FD_CHECK(h, "null handle");
    h->prof.clear();
    h->prof.enabled = false;
    h->prof_requested = enable;
    return 0;
}
double fd_profile_ms(fd_handle *h, const char *family) {
    if (!h) return 0.0;
    h->prof.resolve();
    auto it = h->prof.fam.find(family);
    return it == h->prof.fam.end() ? 0.0 : it->second.ms;
}
int64_t fd_profile_launches(fd_handle *h, const char *family) {
    if (!h) return 0;
    auto it = h->prof.fam.find(family);
    return it == h->prof.fam.end() ? 0 : it->second.launches;
}

static int run_score(fd_handle *h, const float *x, const float *temb_row, float *score, int B, cudaStream_t s) {
    return score_generic(h, x, temb_row, score, B, s);  // the driver dispatches per phase on h->active_path
}

int fd_score(fd_handle *h, const float *x_dev, float t, float *score_dev, int32_t batch, void *stream) {
    FD_CHECK(h && x_dev && score_dev && batch > 0, "fd_score: bad argument");
    FD_CHECK(h->finalized, "fd_score: call fd_finalize_weights first");
    FD_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    FD_TRY(ensure_workspace(h, batch, 1, s));
    float *temb_row = h->ws_temb + (size_t)h->cap_steps * h->cfg.d_model;  // spare row after the per-step table
    FD_TRY(launch_time_embedding_scalar(h, t, temb_row, s));
    return run_score(h, x_dev, temb_row, score_dev, batch, s);
}

static void step_coefficients(const fd_config &c, double t, float *cx, float *d0) {
    if (c.sched_kind == FD_SCHED_VP) {  // sde.py:212-213,228-235
        double beta = c.sched_p0 + t * (c.sched_p1 - c.sched_p0);
        *cx = (float)(-0.5 * beta);
        *d0 = (float)sqrt(beta);
    } else {  // sde.py:142-148
        double r = c.sched_p1 / c.sched_p0;
        double sd = c.sched_p0 * sqrt(2.0 * log(r)) * pow(r, t);
        *cx = 0.f;
        *d0 = (float)sd;
    }
}

int fd_step(fd_handle *h, const float *x_dev, const float *score_dev, const float *z_dev, double t, float step_size, float *out_dev,
            int32_t batch, void *stream) {
    FD_CHECK(h && x_dev && score_dev && z_dev && out_dev && batch > 0, "fd_step: bad argument");
    FD_CHECK(step_size > 0.f, "fd_step: step_size must be positive");  // sde.py:239
    FD_CUDA(cudaSetDevice(h->cfg.device));
    float cx, d0;
    step_coefficients(h->cfg, t, &cx, &d0);
    return launch_sde_step(h, x_dev, score_dev, z_dev, out_dev, batch, cx, d0, step_size, sqrtf(step_size), 0, 0, 0,
                           (cudaStream_t)stream);
}

int fd_ffn_block(fd_handle *h, int32_t layer, float *h_dev, int32_t n_tokens, void *stream) {
    FD_CHECK(h && h_dev && n_tokens > 0, "fd_ffn_block: bad argument");
    FD_CHECK(h->finalized && h->cfg.model_kind == FD_MODEL_TRANSFORMER, "fd_ffn_block: needs a finalized transformer handle");
    FD_CHECK(layer >= 0 && layer < (int)h->tl.size(), "fd_ffn_block: layer %d out of range", layer);
    FD_CUDA(cudaSetDevice(h->cfg.device));
    return ffn_block(h, layer, h_dev, n_tokens, (cudaStream_t)stream);
}

int fd_attention_block(fd_handle *h, int32_t layer, float *h_dev, int32_t batch, void *stream) {
    FD_CHECK(h && h_dev && batch > 0, "fd_attention_block: bad argument");
    FD_CHECK(h->finalized && h->cfg.model_kind == FD_MODEL_TRANSFORMER, "fd_attention_block: needs a finalized transformer handle");
    FD_CHECK(layer >= 0 && layer < (int)h->tl.size(), "fd_attention_block: layer %d out of range", layer);
    FD_CUDA(cudaSetDevice(h->cfg.device));
    FD_TRY(ensure_workspace(h, batch, 1, (cudaStream_t)stream));
    return attention_block(h, layer, h_dev, batch, (cudaStream_t)stream);
}

int fd_encoder_stack(fd_handle *h, float *h_dev, int32_t batch, void *stream) {
    FD_CHECK(h && h_dev && batch > 0, "fd_encoder_stack: bad argument");
    FD_CHECK(h->finalized && h->cfg.model_kind == FD_MODEL_TRANSFORMER, "fd_encoder_stack: needs a finalized transformer handle");
    FD_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    FD_TRY(ensure_workspace(h, batch, 1, s));
    const size_t bytes = (size_t)batch * h->cfg.max_len * h->cfg.d_model * sizeof(float);
    FD_CUDA(cudaMemcpyAsync(h->ws_h, h_dev, bytes, cudaMemcpyDeviceToDevice, s));
    h->himg_primed = 0;  // rows only: the first attention task / kernel gathers its token tile
    FD_TRY(transformer_layers(h, batch, s));
    FD_CUDA(cudaMemcpyAsync(h_dev, h->ws_h, bytes, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int fd_prior(fd_handle *h, const float *z_dev, float *out_dev, int32_t batch, void *stream) {
    FD_CHECK(h && z_dev && out_dev && batch > 0, "fd_prior: bad argument");
    FD_CUDA(cudaSetDevice(h->cfg.device));
    return launch_prior(h, z_dev, out_dev, batch, 0, 0, (cudaStream_t)stream);
}

int fd_normal(fd_handle *h, uint64_t seed, uint64_t first_series, uint32_t draw, float *out_dev, int32_t batch, void *stream) {
    FD_CHECK(h && out_dev && batch > 0, "fd_normal: bad argument");
    FD_CUDA(cudaSetDevice(h->cfg.device));
    return launch_normal(h, out_dev, batch, seed, first_series, draw, (cudaStream_t)stream);
}

int fd_sample(fd_handle *h, int32_t batch, int32_t n_run, const float *timesteps_host, float step_size, uint64_t seed,
              uint64_t first_series, const float *prior_z_dev, const float *noise_dev, float *out_dev, void *stream) {
    FD_CHECK(h && timesteps_host && out_dev && batch > 0 && n_run >= 0, "fd_sample: bad argument");
    FD_CHECK(h->finalized, "fd_sample: call fd_finalize_weights first");
    FD_CHECK(step_size > 0.f, "fd_sample: step_size must be positive");
    FD_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    const fd_config &c = h->cfg;
    const size_t per_batch = (size_t)batch * c.max_len * c.n_channels;
    FD_TRY(ensure_workspace(h, batch, n_run > 0 ? n_run : 1, s));
    if (n_run > 0) {
        FD_CUDA(cudaMemcpyAsync(h->ws_tsteps, timesteps_host, (size_t)n_run * sizeof(float), cudaMemcpyHostToDevice, s));
        FD_TRY(launch_time_embedding(h, h->ws_tsteps, n_run, h->ws_temb, s));
    }
    FD_TRY(launch_prior(h, prior_z_dev, h->ws_x, batch, seed, first_series, s));
    const float sqrt_dt = sqrtf(step_size);
    const int stride = h->prof_requested > 0 ? h->prof_requested : 0;
    if (h->lstm_persistent && lstm_stack_tc_supported(h) && n_run > 0) {
        // LSTM score network: the whole reverse-diffusion loop is ONE launch (a CTA keeps its series on chip for all n_run steps)
        std::vector<float> coef(2 * (size_t)n_run);
        for (int i = 0; i < n_run; ++i) step_coefficients(c, (double)timesteps_host[i], &coef[2 * i], &coef[2 * i + 1]);
        // (pageable source: the copy is staged before cudaMemcpyAsync returns, so the vector may go out of scope)
        FD_CUDA(cudaMemcpyAsync(h->ws_coef, coef.data(), coef.size() * sizeof(float), cudaMemcpyHostToDevice, s));
        h->prof.enabled = stride > 0;
        h->prof.begin("lstm_sampler", s);
        FD_TRY(launch_lstm_sampler(h, h->ws_x, nullptr, h->ws_temb, h->ws_coef, noise_dev, batch, n_run, step_size, sqrt_dt, seed, first_series, s));
        h->prof.end("lstm_sampler", s, 1);
        h->prof.enabled = false;
        FD_CUDA(cudaMemcpyAsync(out_dev, h->ws_x, per_batch * sizeof(float), cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    // Series are independent, so the batch can be cut into independent sub-batches ("lanes", default 2) whose kernels are issued on separate streams: whenever one
    // half's kernel leaves SMs idle (partial last wave: 256 FFN CTAs or 1024 attention CTAs do not divide 148 SMs), the other half's
    // CTAs fill them.  Steps that are being profiled run un-split on the caller's stream so that kernel durations are clean.
    // With the persistent stack kernel ("stack_lanes") a second sub-batch's kernel fills the SMs that the first one's drains: its CTAs
    // become resident as the first kernel's CTAs run out of tasks, and the tensor-bound FFN-only tail of one queue overlaps the
    // exponential-bound attention-only head of the next.  Each lane has its own queue / dependency-counter state (stack_select_slot);
    // the samples are bit-identical to the un-split run.  Measured in one gpurun call (profiles/r02g_ab_stack_lanes.txt): batch 1024 x
    // max_len 252 (cfg 3) 253.5 -> 258.4 (2 lanes) -> 260.4 series/s (3 lanes); batch 256 (cfg 2) 271.5 -> 257.8 -> 206.8: a lane needs a
    // few hundred series for its queue head / tail to amortise, so the default splits only into lanes of >= 340 series.
    const bool stack = stack_supported(h);
    const int lanes_env = stack ? (h->stack_lanes > 0 ? h->stack_lanes : std::max(1, std::min(3, batch / 340))) : h->lanes;
    int nl = (lanes_env >= 2 && h->active_path == 1 && h->attn_fast && batch >= 32) ? lanes_env : 1;
    if (nl > 1 && batch < 16 * nl) nl = 2;
    if (nl > 1 && !h->lane_stream[0]) {
        for (int k = 0; k < FD_MAX_LANES; ++k) FD_CUDA(cudaStreamCreateWithFlags(&h->lane_stream[k], cudaStreamNonBlocking));
        for (int k = 0; k <= FD_MAX_LANES; ++k) FD_CUDA(cudaEventCreateWithFlags(&h->lane_event[k], cudaEventDisableTiming));
    }
    auto lane_lo = [&](int k) { return (int)(((long long)batch * k) / nl); };  // lane k owns series [lane_lo(k), lane_lo(k+1))
    const size_t LC = (size_t)c.max_len * c.n_channels, LD = (size_t)c.max_len * c.d_model;
    struct View { float *x, *score, *hh, *h2, *att, *qkv, *himg, *attimg, *qimg, *kvimg; unsigned *nrm; } base = {
        h->ws_x, h->ws_score, h->ws_h, h->ws_h2, h->ws_att, h->ws_qkv, h->ws_himg, h->ws_attimg, h->ws_qimg, h->ws_kvimg, h->ws_nrm};
    auto set_view = [&](int b0, int lane) {  // the drivers read their workspace pointers from the handle: point them at the half-batch
        const size_t wide = (size_t)c.max_len * (c.model_kind == FD_MODEL_LSTM ? 4 : 3) * c.d_model;
        h->ws_x = base.x + b0 * LC;
        h->ws_score = base.score + b0 * LC;
        h->ws_h = base.hh + b0 * LD;
        h->ws_h2 = base.h2 + b0 * LD;
        h->ws_att = base.att + b0 * LD;
        h->ws_qkv = base.qkv + b0 * wide;
        if (base.attimg) {  // series images are per series; each lane owns whole 256-token tiles of the attention image (disjoint by construction)
            if (base.himg) h->ws_himg = base.himg + (size_t)b0 * (18 * 256 * 4);
            h->ws_attimg = base.attimg + ((size_t)b0 * c.max_len / 256 + lane) * (9 * 256 * 4);
            if (base.qimg) {
                h->ws_qimg = base.qimg + stream_qimg_floats(b0, c.max_len);
                h->ws_kvimg = base.kvimg + stream_kvimg_floats(b0, c.max_len);
                h->ws_nrm = base.nrm + stream_nrm_words(b0);
            }
        }
    };
    const bool fused_boundary = h->fuse_boundary && step_boundary_supported(h);
    int mode = 1;  // 1: everything on `s`; nl: the lanes are in flight
    auto to_mode = [&](int want) -> int {
        if (want == mode) return 0;
        if (want > 1) {  // fork
            FD_CUDA(cudaEventRecord(h->lane_event[FD_MAX_LANES], s));
            for (int k = 0; k < nl; ++k) FD_CUDA(cudaStreamWaitEvent(h->lane_stream[k], h->lane_event[FD_MAX_LANES], 0));
        } else {  // join
            for (int k = 0; k < nl; ++k) {
                FD_CUDA(cudaEventRecord(h->lane_event[k], h->lane_stream[k]));
                FD_CUDA(cudaStreamWaitEvent(s, h->lane_event[k], 0));
            }
        }
        mode = want;
        return 0;
    };
    int rc = 0;
    for (int i = 0; i < n_run && !rc; ++i) {
        const bool prof_step = stride > 0 && (i % stride == 0);
        h->prof.enabled = prof_step;
        float cx, d0;
        step_coefficients(c, (double)timesteps_host[i], &cx, &d0);
        const int want = (nl > 1 && !prof_step) ? nl : 1;
        if ((rc = to_mode(want))) break;
        for (int k = 0; k < want && !rc; ++k) {
            const int b0 = want > 1 ? lane_lo(k) : 0;
            const int nb = want > 1 ? lane_lo(k + 1) - lane_lo(k) : batch;
            cudaStream_t sk = want > 1 ? h->lane_stream[k] : s;
            set_view(b0, want > 1 ? k : 0);
            if (stack) stack_select_slot(h, want > 1 ? k + 1 : 0);
            const float *z = noise_dev ? noise_dev + (size_t)i * per_batch + b0 * LC : nullptr;
            if (fused_boundary) {
                // embed of step 0 here; afterwards the boundary kernel (unembed + scheduler step + embed for the next step) keeps ws_h primed
                if (i == 0) rc = transformer_embed(h, h->ws_x, h->ws_temb, nb, sk);
                if (!rc) rc = transformer_layers(h, nb, sk);
                if (rc) break;
                h->prof.begin("boundary", sk);
                rc = launch_step_boundary(h, h->ws_h, h->ws_x, z, h->ws_temb + (size_t)(i + 1) * c.d_model, nb, cx, d0, step_size, sqrt_dt, seed,
                                          first_series + b0, (uint32_t)(i + 1), i + 1 < n_run, sk);
                h->prof.end("boundary", sk, 1);
                continue;
            }
            rc = run_score(h, h->ws_x, h->ws_temb + (size_t)i * c.d_model, h->ws_score, nb, sk);
            if (rc) break;
            h->prof.begin("sde_step", sk);
            rc = launch_sde_step(h, h->ws_x, h->ws_score, z, h->ws_x, nb, cx, d0, step_size, sqrt_dt, seed, first_series + b0,
                                 (uint32_t)(i + 1), sk);
            h->prof.end("sde_step", sk, 1);
        }
        set_view(0, 0);
    }
    set_view(0, 0);
    if (stack) stack_select_slot(h, 0);
    {   // join the lanes on the error path too: the caller's stream must not run ahead of work that is still in flight on them
        const int rj = to_mode(1);
        if (!rc) rc = rj;
    }
    if (rc) return rc;
    h->prof.enabled = false;
    FD_CUDA(cudaMemcpyAsync(out_dev, h->ws_x, per_batch * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
}

int fd_sample_host(fd_handle *h, int32_t batch, int32_t n_run, const float *timesteps_host, float step_size, uint64_t seed,
                   uint64_t first_series, const float *prior_z_host, const float *noise_host, float *out_host, void *stream) {
    FD_CHECK(h && out_host && batch > 0 && n_run >= 0, "fd_sample_host: bad argument");
    FD_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t per_batch = (size_t)batch * h->cfg.max_len * h->cfg.n_channels;
    const size_t out_bytes = per_batch * sizeof(float);
    size_t noise_bytes = 0;
    if (prior_z_host) noise_bytes += out_bytes;
    if (noise_host) noise_bytes += out_bytes * (size_t)n_run;
    if (noise_bytes > h->stage_noise_bytes) {
        if (h->stage_noise) cudaFree(h->stage_noise);
        h->stage_noise = nullptr;
        FD_CUDA(cudaMalloc((void **)&h->stage_noise, noise_bytes));
        h->stage_noise_bytes = noise_bytes;
    }
    if (out_bytes > h->stage_out_bytes) {
        if (h->stage_out) cudaFree(h->stage_out);
        h->stage_out = nullptr;
        FD_CUDA(cudaMalloc((void **)&h->stage_out, out_bytes));
        h->stage_out_bytes = out_bytes;
    }
    float *pz = nullptr, *nz = nullptr;
    float *cursor = h->stage_noise;
    if (prior_z_host) {
        pz = cursor;
        cursor += per_batch;
        FD_CUDA(cudaMemcpyAsync(pz, prior_z_host, out_bytes, cudaMemcpyHostToDevice, s));
    }
    if (noise_host) {
        nz = cursor;
        FD_CUDA(cudaMemcpyAsync(nz, noise_host, out_bytes * (size_t)n_run, cudaMemcpyHostToDevice, s));
    }
    FD_TRY(fd_sample(h, batch, n_run, timesteps_host, step_size, seed, first_series, pz, nz, h->stage_out, s));
    FD_CUDA(cudaMemcpyAsync(out_host, h->stage_out, out_bytes, cudaMemcpyDeviceToHost, s));
    FD_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int fd_dft(const float *x_dev, float *out_dev, int32_t batch, int32_t max_len, int32_t n_channels, int32_t device, void *stream) {
    FD_CHECK(x_dev && out_dev && batch > 0 && max_len > 0 && n_channels > 0, "fd_dft: bad argument");
    FD_CUDA(cudaSetDevice(device));
    return launch_dft(x_dev, out_dev, batch, max_len, n_channels, nullptr, nullptr, false, (cudaStream_t)stream);
}

int fd_idft(const float *x_dev, float *out_dev, int32_t batch, int32_t max_len, int32_t n_channels, const float *mean_dev,
            const float *std_dev, int32_t device, void *stream) {
    FD_CHECK(x_dev && out_dev && batch > 0 && max_len > 0 && n_channels > 0, "fd_idft: bad argument");
    FD_CHECK((mean_dev == nullptr) == (std_dev == nullptr), "fd_idft: mean and std must be given together");
    FD_CUDA(cudaSetDevice(device));
    return launch_dft(x_dev, out_dev, batch, max_len, n_channels, mean_dev, std_dev, true, (cudaStream_t)stream);
}

int fd_spectral_density(const float *x_dev, float *out_dev, float *scratch_dev, int32_t batch, int32_t max_len, int32_t n_channels,
                        int32_t apply_dft, int32_t device, void *stream) {
    FD_CHECK(x_dev && out_dev && batch > 0 && max_len > 0 && n_channels > 0, "fd_spectral_density: bad argument");
    FD_CHECK(!apply_dft || scratch_dev, "fd_spectral_density: apply_dft needs a (batch, max_len, n_channels) scratch buffer");
    FD_CUDA(cudaSetDevice(device));
    const float *packed = x_dev;
    if (apply_dft) {
        FD_TRY(launch_dft(x_dev, scratch_dev, batch, max_len, n_channels, nullptr, nullptr, false, (cudaStream_t)stream));
        packed = scratch_dev;
    }
    return launch_spectral_density(packed, out_dev, batch, max_len, n_channels, (cudaStream_t)stream);
}

int fd_wasserstein(const float *x_dev, const float *y_dev, const double *dirs_dev, int32_t n, int32_t m, int32_t d, int32_t n_dirs,
                   int32_t standardise, double *out_dev, int32_t device, void *stream) {
    FD_CHECK(x_dev && y_dev && out_dev && n > 0 && m > 0 && d > 0 && n_dirs > 0, "fd_wasserstein: bad argument");
    FD_CHECK(dirs_dev || n_dirs == d, "fd_wasserstein: marginal distances (no directions) need n_dirs == d");
    FD_CHECK(n_dirs <= 65535, "fd_wasserstein: at most 65535 directions per call");
    FD_CHECK((long long)n * m < (1ll << 62) && n <= (1 << 28) && m <= (1 << 28), "fd_wasserstein: sample sets too large");
    FD_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    float *work = nullptr;
    FD_CUDA(cudaMallocAsync((void **)&work, wasserstein_work_bytes(n, m, n_dirs), s));
    const int rc = launch_wasserstein(x_dev, y_dev, dirs_dev, n, m, d, n_dirs, standardise, work, out_dev, s);
    cudaFreeAsync(work, s);
    return rc;
}

int fd_feature_stats(const float *x_dev, float *mean_dev, float *std_dev, int64_t n, int32_t n_features, int32_t device, void *stream) {
    FD_CHECK(x_dev && mean_dev && std_dev && n > 0 && n_features > 0, "fd_feature_stats: bad argument");
    FD_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    double *work = nullptr;
    FD_CUDA(cudaMallocAsync((void **)&work, feature_stats_work_bytes(n_features), s));
    const int rc = launch_feature_stats(x_dev, mean_dev, std_dev, n, n_features, work, s);
    cudaFreeAsync(work, s);
    return rc;
}

int fd_standardise(const float *x_dev, const float *mean_dev, const float *std_dev, float *out_dev, int64_t n, int32_t n_features, int32_t inverse,
                   int32_t device, void *stream) {
    FD_CHECK(x_dev && mean_dev && std_dev && out_dev && n > 0 && n_features > 0, "fd_standardise: bad argument");
    FD_CUDA(cudaSetDevice(device));
    return launch_standardise(x_dev, mean_dev, std_dev, out_dev, n, n_features, inverse, (cudaStream_t)stream);
}

}  // extern "C"
