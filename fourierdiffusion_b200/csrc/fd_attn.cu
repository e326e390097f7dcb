// TF32 tensor-core path, part 2: the token-wise projections and multi-head attention of one encoder layer
// (nn.TransformerEncoderLayer reached from score_models.py:87; math spec SURVEY.md appendix A.5), d_model = 72, max_len <= 256.
//
//   linear72_kernel<QKV>   qkv = h · Win^T + bin for a 128-token tile (tcgen05 kind::tf32, M=128, N=240, K=72); the epilogue writes
//                          q (pre-scaled by log2(e)/sqrt(dh)), k and v^T straight into the per-(series, head) shared-memory IMAGES
//                          the attention kernel's UMMA descriptors expect, so attention stages a head with ONE bulk copy.
//   attention_kernel       per (series, head, 128-query tile):  S = Q K^T (one tcgen05.mma, N = keys) -> TMEM; softmax rows in
//                          registers (thread = query row: tcgen05.ld, max, ex2, tcgen05.st in place); O = P V (tcgen05.mma, A = P from
//                          TMEM, B = V^T image, N = 16) with a ones-column in V^T so column 6 of O is the softmax denominator.
//   linear72_kernel<OUT>   h <- LN1(h + att · Wo^T + bo) (M=128, N=80, K=72), LayerNorm in the epilogue (thread = token row).
#include <math.h>

#include "fd_common.cuh"
#include "fd_tc.cuh"

namespace fd {

using namespace tc;

namespace att {
constexpr int D = 72, KC = 18, H = 12, DH = 6;
constexpr int LP = 256;                          // rows of a head image (max_len <= 256)
constexpr int IMG_Q = 0, IMG_K = 2 * LP * 4, IMG_V = 4 * LP * 4;  // float offsets inside a head image
constexpr int VROWS = 8;                         // v^T image rows: d = 0..5, the ones-row (6) and a zero row; the UMMA N=16 operand reads
                                                 // rows 8..15 through SBO = 0, i.e. as copies of rows 0..7 (those output columns are ignored)
constexpr int IMG_FLOATS = IMG_V + (LP / 4) * VROWS * 4;          // 6144 floats = 24576 B
constexpr int IMG_BYTES = IMG_FLOATS * 4;
constexpr int TMT = 128;                         // tokens per CTA of the linear kernels
constexpr int NP_QKV = 240;                      // q at columns 0..71, k at 80..151, v at 160..231 (sections 16-aligned)
constexpr int NP_OUT = 80;
constexpr int X_BYTES = KC * TMT * 16;           // 36864
}  // namespace att

enum { LIN_QKV = 0, LIN_OUT = 1 };

// ---- weight images -----------------------------------------------------------------------------------------------------------
// out[(kc, n, j)] = tf32(W[row(n)][kc*4 + j]) for the [kc][NP][4] UMMA image; QKV: n in section s (stride 80) maps to W row s*72 + n%80.
__global__ void pack_linear_weights_kernel(const float *__restrict__ w, float *__restrict__ out, int NP, int mode) {
    using namespace att;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < KC * NP * 4; i += gridDim.x * blockDim.x) {
        int j = i % 4, n = (i / 4) % NP, kc = i / (4 * NP);
        int row = -1;
        if (mode == LIN_QKV) {
            int s = n / 80, r = n % 80;
            if (r < D) row = s * D + r;
        } else if (n < D) {
            row = n;
        }
        out[i] = row >= 0 ? __uint_as_float(f32_to_tf32(w[(size_t)row * D + kc * 4 + j])) : 0.f;
    }
}

// ones-column of V^T (image row d = 6) for the valid keys of every (series, head) image; everything else stays zero
__global__ void init_qkv_images_kernel(float *__restrict__ img, int n_images, int L) {
    using namespace att;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n_images * L; i += (long long)gridDim.x * blockDim.x) {
        int im = (int)(i / L), pos = (int)(i % L);
        img[(size_t)im * IMG_FLOATS + IMG_V + ((pos / 4) * VROWS + 6) * 4 + (pos % 4)] = 1.0f;
    }
}

__device__ __forceinline__ void load_row72(uint32_t taddr, float (&y)[72]) {
    uint32_t v[32];
    tmem_ld32(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = __uint_as_float(v[j]);
    tmem_ld32(taddr + 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) y[32 + j] = __uint_as_float(v[j]);
    uint32_t u[8];
    tmem_ld8(taddr + 64, u);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) y[64 + j] = __uint_as_float(u[j]);
}

// ---- token-wise linear layers on the tensor cores ---------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128, 2)
linear72_kernel(const float *x_in, const float *__restrict__ wimg, const float *__restrict__ bias, float *h_io,
                const float *__restrict__ ln_w, const float *__restrict__ ln_b, float *__restrict__ qkv_img, int M, int L, float qscale) {
    using namespace att;
    constexpr int NP = MODE == LIN_QKV ? NP_QKV : NP_OUT;
    constexpr int W_BYTES = KC * NP * 16;
    constexpr int TCOLS = MODE == LIN_QKV ? 256 : 128;
    extern __shared__ __align__(1024) uint8_t smem[];
    float *Xs = reinterpret_cast<float *>(smem);
    uint8_t *Ws = smem + X_BYTES;
    const uint32_t bar_w = smem_u32(smem + X_BYTES + W_BYTES), bar_mma = bar_w + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + X_BYTES + W_BYTES + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * TMT;

    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(bar_w, W_BYTES);
        bulk_g2s(smem_u32(Ws), wimg, W_BYTES, bar_w);
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), TCOLS);
    }
    for (int idx = tid; idx < KC * TMT; idx += 128) {
        int row = idx % TMT, kc = idx / TMT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + row < M) v = *reinterpret_cast<const float4 *>(x_in + (size_t)(m0 + row) * D + kc * 4);
        reinterpret_cast<uint4 *>(Xs)[idx] = make_uint4(f32_to_tf32(v.x), f32_to_tf32(v.y), f32_to_tf32(v.z), f32_to_tf32(v.w));
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        mbar_wait(bar_w, 0);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t idesc = make_idesc_tf32(128, NP);
            const uint64_t a0 = make_smem_desc(smem_u32(Xs), TMT * 16, 128), b0 = make_smem_desc(smem_u32(Ws), NP * 16, 128);
#pragma unroll
            for (int ks = 0; ks < D / 8; ++ks)
                mma_tf32_ss(tmem, a0 + (uint64_t)(ks * 2 * (TMT * 16) >> 4), b0 + (uint64_t)(ks * 2 * (NP * 16) >> 4), idesc, ks > 0);
            mma_commit(bar_mma);
        }
        __syncwarp();
    }
    mbar_wait(bar_mma, 0);
    tc_fence_after();

    const int token = m0 + 32 * warp + lane;
    const uint32_t trow = tmem + ((uint32_t)(32 * warp) << 16);
    float y[72];
    if (MODE == LIN_QKV) {
        const int b = token / L, pos = token % L;
        float *img0 = qkv_img + (size_t)b * H * IMG_FLOATS;
#pragma unroll 1
        for (int sec = 0; sec < 3; ++sec) {
            load_row72(trow + sec * 80, y);
            if (token < M) {
                const float sc = sec == 0 ? qscale : 1.0f;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + sec * D) + k);
                    y[4 * k + 0] = (y[4 * k + 0] + bb.x) * sc;
                    y[4 * k + 1] = (y[4 * k + 1] + bb.y) * sc;
                    y[4 * k + 2] = (y[4 * k + 2] + bb.z) * sc;
                    y[4 * k + 3] = (y[4 * k + 3] + bb.w) * sc;
                }
                if (sec < 2) {  // q / k rows of the K-major image [kc][256][4]
#pragma unroll
                    for (int hh = 0; hh < H; ++hh) {
                        float *dst = img0 + (size_t)hh * IMG_FLOATS + (sec == 0 ? IMG_Q : IMG_K) + pos * 4;
                        uint4 lo = make_uint4(f32_to_tf32(y[6 * hh + 0]), f32_to_tf32(y[6 * hh + 1]), f32_to_tf32(y[6 * hh + 2]),
                                              f32_to_tf32(y[6 * hh + 3]));
                        uint4 hi = make_uint4(f32_to_tf32(y[6 * hh + 4]), f32_to_tf32(y[6 * hh + 5]), 0u, 0u);
                        *reinterpret_cast<uint4 *>(dst) = lo;
                        *reinterpret_cast<uint4 *>(dst + LP * 4) = hi;
                    }
                } else {  // v^T image [key/4][8][4]
#pragma unroll
                    for (int hh = 0; hh < H; ++hh) {
                        float *dst = img0 + (size_t)hh * IMG_FLOATS + IMG_V + (pos / 4) * (VROWS * 4) + (pos % 4);
#pragma unroll
                        for (int d = 0; d < DH; ++d) dst[d * 4] = __uint_as_float(f32_to_tf32(y[6 * hh + d]));
                    }
                }
            }
        }
    } else {
        load_row72(trow, y);
        if (token < M) {
            float *hrow = h_io + (size_t)token * D;
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                float4 r = reinterpret_cast<const float4 *>(hrow)[k];
                float4 bb = __ldg(reinterpret_cast<const float4 *>(bias) + k);
                y[4 * k + 0] += r.x + bb.x;
                y[4 * k + 1] += r.y + bb.y;
                y[4 * k + 2] += r.z + bb.z;
                y[4 * k + 3] += r.w + bb.w;
                sum += y[4 * k + 0] + y[4 * k + 1] + y[4 * k + 2] + y[4 * k + 3];
            }
            const float mean = sum * (1.0f / D);
            float var = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                float d = y[j] - mean;
                var = fmaf(d, d, var);
            }
            const float rstd = 1.0f / sqrtf(var * (1.0f / D) + 1e-5f);
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                float4 w = __ldg(reinterpret_cast<const float4 *>(ln_w) + k);
                float4 bb = __ldg(reinterpret_cast<const float4 *>(ln_b) + k);
                float4 o;
                o.x = (y[4 * k + 0] - mean) * rstd * w.x + bb.x;
                o.y = (y[4 * k + 1] - mean) * rstd * w.y + bb.y;
                o.z = (y[4 * k + 2] - mean) * rstd * w.z + bb.z;
                o.w = (y[4 * k + 3] - mean) * rstd * w.w + bb.w;
                reinterpret_cast<float4 *>(hrow)[k] = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

// ---- attention ------------------------------------------------------------------------------------------------------------------------
namespace att {
constexpr int ATT_THREADS = 192;
constexpr int COL_S = 0, COL_O = 256;
}  // namespace att

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Row softmax of one 128-query tile straight out of TMEM: thread = query row, S columns = keys (already in log2 units).
// Pass 1 finds the row maximum, pass 2 writes P = 2^(s - max) (tf32-rounded) back in place and signals the MMA warp per 64-key
// quarter so the P·V MMAs of a quarter overlap the exponentials of the next.  128 columns are fetched per tcgen05.wait::ld.
template <bool FULL>
__device__ __forceinline__ void softmax_rows(uint32_t tS, int L, uint32_t p_ready0) {
    const int nq = (L + 63) / 64;
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll 1
    for (int g = 0; g < nq; ++g) {  // one 64-key quarter per tcgen05.wait::ld
        uint32_t v[2][32];
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (FULL || (g * 64 + i * 32 < L)) tmem_ld32(tS + g * 64 + i * 32, v[i]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (FULL || (g * 64 + i * 32 < L)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int col = g * 64 + i * 32 + j;
                    if (FULL || col + 3 < L) {
                        m0 = fmaxf(m0, __uint_as_float(v[i][j]));
                        m1 = fmaxf(m1, __uint_as_float(v[i][j + 1]));
                        m2 = fmaxf(m2, __uint_as_float(v[i][j + 2]));
                        m3 = fmaxf(m3, __uint_as_float(v[i][j + 3]));
                    } else {
                        if (col < L) m0 = fmaxf(m0, __uint_as_float(v[i][j]));
                        if (col + 1 < L) m1 = fmaxf(m1, __uint_as_float(v[i][j + 1]));
                        if (col + 2 < L) m2 = fmaxf(m2, __uint_as_float(v[i][j + 2]));
                    }
                }
            }
        }
    }
    const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
#pragma unroll 1
    for (int g = 0; g < nq; ++g) {
        uint32_t v[2][32];
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (FULL || (g * 64 + i * 32 < L)) tmem_ld32(tS + g * 64 + i * 32, v[i]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (FULL || (g * 64 + i * 32 < L)) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const uint32_t bits = __float_as_uint(ex2_approx(__uint_as_float(v[i][j]) - m)) + 0x1000u;  // tf32 rounding
                    v[i][j] = (FULL || g * 64 + i * 32 + j < L) ? bits : 0u;
                }
                tmem_st32(tS + g * 64 + i * 32, v[i]);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_ready0 + 8u * g);
    }
}

__global__ void __launch_bounds__(att::ATT_THREADS, 1)
attention_kernel(const float *__restrict__ qkv_img, float *__restrict__ att_out, int L, int heads_per_cta) {
    using namespace att;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x, head0 = blockIdx.y * heads_per_cta;
    const uint32_t img_smem = smem_u32(smem);
    const uint32_t bar0 = smem_u32(smem + 2 * IMG_BYTES);
    auto QKV_FULL = [&](int i) { return bar0 + 8u * i; };
    auto QKV_EMPTY = [&](int i) { return bar0 + 8u * (2 + i); };
    const uint32_t S_FULL = bar0 + 32, O_FULL = bar0 + 40, P_READY0 = bar0 + 48;  // P_READY0 + 8*quarter
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 2 * IMG_BYTES + 96);
    const int NT = (L + 127) / 128;
    const uint8_t *src = reinterpret_cast<const uint8_t *>(qkv_img) + ((size_t)b * H + head0) * IMG_BYTES;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(QKV_FULL(i), 1);
            mbar_init(QKV_EMPTY(i), 1);
        }
        mbar_init(S_FULL, 1);
        mbar_init(O_FULL, 1);
        for (int i = 0; i < 4; ++i) mbar_init(P_READY0 + 8u * i, 128);
        mbar_fence_init();
    }
    if (warp == 5) {
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 5) {
        // ===== producer: one bulk copy per head image =====
        if (lane == 0) {
            for (int hh = 0; hh < heads_per_cta; ++hh) {
                const int buf = hh & 1, use = hh >> 1;
                mbar_wait(QKV_EMPTY(buf), (use & 1) ^ 1);
                mbar_arrive_expect_tx(QKV_FULL(buf), IMG_BYTES);
                bulk_g2s(img_smem + buf * IMG_BYTES, src + (size_t)hh * IMG_BYTES, IMG_BYTES, QKV_FULL(buf));
            }
        }
    } else if (warp == 4) {
        // ===== MMA issuer (warp-uniform loop, the elected lane issues) =====
        const int NK = ((L + 15) / 16) * 16;
        const uint32_t idesc_s = make_idesc_tf32(128, NK), idesc_o = make_idesc_tf32(128, 16);
        const int ksteps = (L + 7) / 8, nq = (L + 63) / 64;
        const uint32_t leader = elect_one() ? 1u : 0u;
        int task = 0;
        for (int hh = 0; hh < heads_per_cta; ++hh) {
            const int buf = hh & 1;
            mbar_wait(QKV_FULL(buf), (hh >> 1) & 1);
            tc_fence_after();
            const uint32_t base = img_smem + buf * IMG_BYTES;
            const uint64_t kd = make_smem_desc(base + IMG_K * 4, LP * 16, 128);
            const uint64_t vd = make_smem_desc(base + IMG_V * 4, VROWS * 16, 0);  // SBO 0: rows 8..15 alias rows 0..7
            for (int t = 0; t < NT; ++t, ++task) {
                const uint64_t qd = make_smem_desc(base + IMG_Q * 4 + t * 128 * 16, LP * 16, 128);
                mma_tf32_ss_if(leader, tmem + COL_S, qd, kd, idesc_s, 0);
                mma_commit_if(leader, S_FULL);
                for (int qt = 0; qt < nq; ++qt) {
                    mbar_wait(P_READY0 + 8u * qt, task & 1);
                    tc_fence_after();
#pragma unroll
                    for (int k8 = 0; k8 < 8; ++k8) {
                        const int ks = qt * 8 + k8;
                        if (ks < ksteps)
                            mma_tf32_ts_if(leader, tmem + COL_O, tmem + COL_S + ks * 8, vd + (uint64_t)(ks * (2 * VROWS * 16 >> 4)), idesc_o,
                                           ks > 0);
                    }
                }
                mma_commit_if(leader, O_FULL);
                if (t == NT - 1) mma_commit_if(leader, QKV_EMPTY(buf));
            }
        }
    } else {
        // ===== softmax warps: thread = query row =====
        const uint32_t trow = tmem + ((uint32_t)(32 * warp) << 16);
        int task = 0;
        for (int hh = 0; hh < heads_per_cta; ++hh) {
            for (int t = 0; t < NT; ++t, ++task) {
                mbar_wait(S_FULL, task & 1);
                tc_fence_after();
                if (L == LP)
                    softmax_rows<true>(trow + COL_S, L, P_READY0);
                else
                    softmax_rows<false>(trow + COL_S, L, P_READY0);
                // O epilogue: columns 0..5 = sum_k P V, column 6 = sum_k P
                mbar_wait(O_FULL, task & 1);
                tc_fence_after();
                uint32_t o[8];
                tmem_ld8(trow + COL_O, o);
                tmem_ld_wait();
                const int q = t * 128 + 32 * warp + lane;
                if (q < L) {
                    const float inv = 1.0f / __uint_as_float(o[6]);
                    float *dst = att_out + ((size_t)b * L + q) * D + (head0 + hh) * DH;
                    reinterpret_cast<float2 *>(dst)[0] = make_float2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                    reinterpret_cast<float2 *>(dst)[1] = make_float2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                    reinterpret_cast<float2 *>(dst)[2] = make_float2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem, 512);
}

// ---- host side ------------------------------------------------------------------------------------------------------------------------
namespace att {
constexpr int SMEM_QKV = X_BYTES + KC * NP_QKV * 16 + 64;
constexpr int SMEM_OUT = X_BYTES + KC * NP_OUT * 16 + 64;
constexpr int SMEM_ATT = 2 * IMG_BYTES + 128;
}  // namespace att

int attn_path_supported(const fd_config &c) {
    return c.model_kind == FD_MODEL_TRANSFORMER && c.d_model == att::D && c.n_head == att::H && c.max_len <= att::LP && c.max_len >= 8;
}

int attn_finalize(fd_handle *h) {
    using namespace att;
    for (auto &w : h->tl) {
        float *a = nullptr, *b = nullptr;
        FD_CUDA(cudaMalloc((void **)&a, (size_t)KC * NP_QKV * 16));
        FD_CUDA(cudaMalloc((void **)&b, (size_t)KC * NP_OUT * 16));
        h->owned.push_back(a);
        h->owned.push_back(b);
        pack_linear_weights_kernel<<<64, 256>>>(w.in_w, a, NP_QKV, LIN_QKV);
        pack_linear_weights_kernel<<<64, 256>>>(w.out_w, b, NP_OUT, LIN_OUT);
        FD_CUDA(cudaGetLastError());
        w.in_pack = a;
        w.out_pack = b;
    }
    // the packed in_proj bias in section order is just in_b (q | k | v), used as is
    FD_CUDA(cudaDeviceSynchronize());
    FD_CUDA(cudaFuncSetAttribute(linear72_kernel<LIN_QKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_QKV));
    FD_CUDA(cudaFuncSetAttribute(linear72_kernel<LIN_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OUT));
    FD_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ATT));
    return 0;
}

// (re)allocate the per-(series, head) q/k/v^T images for `batch` series and set their ones-columns
int attn_ensure_images(fd_handle *h, int batch, cudaStream_t s) {
    using namespace att;
    if (batch <= h->img_batch) return 0;
    if (h->qkv_img) FD_CUDA(cudaFree(h->qkv_img));
    h->qkv_img = nullptr;
    const size_t n = (size_t)batch * H * IMG_FLOATS;
    FD_CUDA(cudaMalloc((void **)&h->qkv_img, n * sizeof(float)));
    FD_CUDA(cudaMemsetAsync(h->qkv_img, 0, n * sizeof(float), s));
    init_qkv_images_kernel<<<296, 256, 0, s>>>(h->qkv_img, batch * H, h->cfg.max_len);
    FD_CUDA(cudaGetLastError());
    h->img_batch = batch;
    return 0;
}

#define FD_KLAUNCH_OK(name)                                                                   \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        FD_CHECK(_e == cudaSuccess, name " launch failed: %s", cudaGetErrorString(_e));       \
        h->launches += 1;                                                                     \
        g_global_launches += 1;                                                               \
    } while (0)

// qkv images <- in_proj(h)
int launch_qkv_fast(fd_handle *h, int layer, const float *hbuf, int B, cudaStream_t s) {
    using namespace att;
    const int L = h->cfg.max_len, M = B * L;
    const TransformerLayerW &w = h->tl[layer];
    const float qscale = (float)(1.4426950408889634 / sqrt((double)DH));
    linear72_kernel<LIN_QKV><<<(M + TMT - 1) / TMT, 128, SMEM_QKV, s>>>(hbuf, w.in_pack, w.in_b, nullptr, nullptr, nullptr, h->qkv_img, M, L,
                                                                         qscale);
    FD_KLAUNCH_OK("linear72_kernel<QKV>");
    return 0;
}

// att_out <- softmax(q k^T / sqrt(dh)) v per head, from the images
int launch_attention_fast(fd_handle *h, float *att_out, int B, cudaStream_t s) {
    using namespace att;
    const int hpc = 3;
    dim3 grid(B, H / hpc);
    attention_kernel<<<grid, ATT_THREADS, SMEM_ATT, s>>>(h->qkv_img, att_out, h->cfg.max_len, hpc);
    FD_KLAUNCH_OK("attention_kernel");
    return 0;
}

// h <- LN1(h + out_proj(att))
int launch_outproj_ln_fast(fd_handle *h, int layer, const float *att_in, float *hbuf, int B, cudaStream_t s) {
    using namespace att;
    const int L = h->cfg.max_len, M = B * L;
    const TransformerLayerW &w = h->tl[layer];
    linear72_kernel<LIN_OUT><<<(M + TMT - 1) / TMT, 128, SMEM_OUT, s>>>(att_in, w.out_pack, w.out_b, hbuf, w.n1_w, w.n1_b, nullptr, M, L, 1.0f);
    FD_KLAUNCH_OK("linear72_kernel<OUT>");
    return 0;
}

}  // namespace fd
