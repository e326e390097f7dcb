// Tensor-core path of the transformer score network (FD_MATH_TF32), specialised for d_model = 72.
//
// Kernel 1 — fused FFN + residual + LayerNorm2 (84 % of the score network's FLOPs at L=256):
//     h <- LN2( h + W2 relu(W1 h + b1) + b2 )                (nn.TransformerEncoderLayer, score_models.py:57-62)
// The GEMM operands are fp16 (tcgen05 kind::f16, fp32 accumulation): fp16 carries the same 11 significant bits as tf32 — the precision
// the reference itself selects on CUDA (cmd/sample.py:23-24) — at twice the tensor-pipe rate and half the shared-memory bytes; its
// narrower exponent is harmless here because every operand is a LayerNorm output, a weight, or a ReLU of their product (conversions
// saturate at +-65504 instead of producing inf).
// One CTA owns 256 tokens (two M=128 UMMA tiles).  The 2048-wide hidden activation never leaves the SM (TMEM): per 64-unit chunk
//     GEMM1  H[128x64]  = [X | 1 1][128x80] · [W1c | b1_hi b1_lo]^T   A and B from shared memory, D in TMEM (double-buffered);
//                                                                      K = 72 is padded to 80 and the spare k-slots carry the bias
//                                                                      (split in two fp16 terms, so it keeps ~22 bits)
//     epi    H <- fp16x2(relu(H))                   tcgen05.ld -> one cvt.rn.relu.satfinite.f16x2 per pair -> tcgen05.st, packed in place
//     GEMM2  Y[128x80] += H[128x64] · W2c^T         A from TMEM (packed fp16), B from shared memory (N padded 72 -> 80)
// and the two token tiles are driven by two independent issuer warps so the tensor pipe works on one tile while the other's epilogue runs.
// Weight chunks (pre-packed on the device into the exact shared-memory image the UMMA descriptors expect) stream from L2 through a
// 4-stage ring of bulk async copies (TMA engine) signalled by mbarriers.  The CTA's token rows stay in a shared-memory slab in fp32 for
// the whole kernel: residual in, LayerNorm1 output (LayerNorm2's residual), result out — nothing but the result goes back to global.
// Warp roles: warp 0 = weight producer (+ TMEM alloc), warps 1-2 = MMA issuers (tile 0 / 1), warps 3-6 / 7-10 = epilogue of tile 0 / 1.
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "fd_common.cuh"
#include "fd_tc.cuh"

namespace fd {

using namespace tc;

namespace fast {
constexpr int D = 72;                            // d_model of the specialised path
constexpr int KC = D / 4;                        // float4 groups of an fp32 token row
constexpr int KP = 80;                           // K of GEMM1 / out-proj padded to a multiple of 16; k = 72, 73 carry the bias
constexpr int KC8 = KP / 8;                      // 16-byte k-chunks (8 halfs) of an operand row
constexpr int TM = 256;                          // tokens per CTA (two M=128 UMMA tiles)
constexpr int NC = 64;                           // hidden units per chunk
constexpr int NY = 80;                           // padded N of GEMM2 (UMMA M=128 needs N % 16 == 0)
constexpr int STAGES = 4;
constexpr int W1_BYTES = KC8 * NC * 16;          // 10240: image [kc][64 rows][8 halfs]
constexpr int W2_BYTES = (NC / 8) * NY * 16;     // 10240: image [kc][80 rows][8 halfs]
constexpr int STAGE_BYTES = W1_BYTES + W2_BYTES; // 20480
constexpr int X_BYTES = KC8 * TM * 16;           // 40960: image [kc][256 rows][8 halfs]
constexpr int ATT_TILE_BYTES = (KC8 - 1) * TM * 16;  // 36864: the 9 data k-chunks of a tile, as the attention kernel writes them
constexpr int THREADS = 352;                     // warp 0 producer, 1-2 MMA issuers (tile 0/1), 3-6 / 7-10 epilogue (tile 0/1)
// TMEM columns: per tile two hidden-chunk buffers H[t][b] (D of GEMM1, A of GEMM2) and the output accumulator Y[t]
constexpr int COL_H = 0;                         // H[t][b] at COL_H + (2*t + b) * NC
constexpr int COL_Y0 = 256, COL_Y1 = 336;
constexpr int TMEM_COLS = 512;
constexpr int RS = 76;                           // LayerNorm slab row stride in floats (16-byte aligned, conflict-free 128-bit row accesses)
constexpr int SLAB_BYTES = 8 * 32 * RS * 4;      // 77824: one 32-row slab per epilogue warp
constexpr int OFF_X = 0;
constexpr int OFF_W = OFF_X + X_BYTES;
constexpr int OFF_SLAB = OFF_W + STAGES * STAGE_BYTES;   // fp32 rows of the CTA's tokens: residual in, LN1 output (= LN2's residual), result out
constexpr int WO_BYTES = KC8 * NY * 16;          // 12800: out_proj image [kc][80][8 halfs] (fused out-proj + LN1 prologue)
constexpr int OFF_WO = OFF_SLAB + SLAB_BYTES;
constexpr int OFF_PAR = OFF_WO + WO_BYTES;       // bo | ln1_w | ln1_b | b2 | ln2_w | ln2_b, 72 floats each
constexpr int OFF_BAR = OFF_PAR + 6 * D * 4;
constexpr int OFF_TMEM = OFF_BAR + 32 * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
}  // namespace fast

// ---- weight packing: fp32 (ff, D) / (D, ff) row-major -> per-chunk fp16 UMMA images -------------------------------------------------
__global__ void pack_ffn_weights_kernel(const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                                        __half *__restrict__ out, int ff) {
    using namespace fast;
    const int per_chunk = STAGE_BYTES / 2;
    const int n_chunks = ff / NC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n_chunks * per_chunk;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i / per_chunk), e = (int)(i % per_chunk);
        float v = 0.f;
        if (e < W1_BYTES / 2) {  // (kc, r, j): W1[c*64 + r][kc*8 + j]; k = 72 / 73: the two fp16 terms of b1[c*64 + r]
            int j = e % 8, r = (e / 8) % NC, k = (e / (8 * NC)) * 8 + j;
            if (k < D) {
                v = w1[(size_t)(c * NC + r) * D + k];
            } else if (k == D || k == D + 1) {
                const float b = b1[c * NC + r];
                const float hi = __half2float(__float2half_rn(b));
                v = k == D ? hi : b - hi;
            }
        } else {  // (kc, n, j): W2[n][c*64 + kc*8 + j], rows 72..79 are zero padding
            int e2 = e - W1_BYTES / 2;
            int j = e2 % 8, n = (e2 / 8) % NY, kc = e2 / (8 * NY);
            if (n < D) v = w2[(size_t)n * ff + c * NC + kc * 8 + j];
        }
        out[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    }
}

// out_proj weight (D, D) row-major -> fp16 image [kc][80 rows][8 halfs] (rows / k beyond 72 zero)
__global__ void pack_outproj16_kernel(const float *__restrict__ wo, __half *__restrict__ out) {
    using namespace fast;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < WO_BYTES / 2; i += gridDim.x * blockDim.x) {
        int j = i % 8, n = (i / 8) % NY, k = (i / (8 * NY)) * 8 + j;
        float v = (n < D && k < D) ? wo[(size_t)n * D + k] : 0.f;
        out[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    }
}

// 8 consecutive fp32 values -> one 16-byte k-chunk of an fp16 operand row (element k in the low half of pair k/2)
__device__ __forceinline__ uint4 pack8_f16(float4 a, float4 b) {
    return make_uint4(pack_f16x2_sat(a.y, a.x), pack_f16x2_sat(a.w, a.z), pack_f16x2_sat(b.y, b.x), pack_f16x2_sat(b.w, b.z));
}

// One token row through residual + bias + LayerNorm on packed fp32 pairs (FADD2 / FFMA2: the epilogues are bound by the FMA pipe's issue
// rate, two lanes per instruction halve it).  y2 = the 72 accumulator columns as 36 pairs; row = the thread's fp32 slab row (residual in,
// normalised row out); bias / w / b = per-column parameters in shared memory.  Two-pass mean / variance like torch's layer_norm.
// ADD = false: y2 already contains residual and bias (the accumulator was initialised with them).
template <bool ADD = true>
__device__ __forceinline__ void residual_layernorm_row(uint64_t (&y2)[36], float *row, const float *bias, const float *w, const float *b) {
    using namespace fast;
    uint64_t sum2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        if (ADD) {
            const ulonglong2 r = *reinterpret_cast<const ulonglong2 *>(row + k * 4);
            const ulonglong2 bb = *reinterpret_cast<const ulonglong2 *>(bias + k * 4);
            y2[2 * k] = f2_add(y2[2 * k], f2_add(r.x, bb.x));
            y2[2 * k + 1] = f2_add(y2[2 * k + 1], f2_add(r.y, bb.y));
        }
        sum2 = f2_add(sum2, f2_add(y2[2 * k], y2[2 * k + 1]));
    }
    float s0, s1;
    f2_unpack(sum2, s0, s1);
    const float mean = (s0 + s1) * (1.0f / D);
    const uint64_t mean2 = f2_pack(mean, mean);
    uint64_t var2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 36; ++i) {
        y2[i] = f2_sub(y2[i], mean2);
        var2 = f2_fma(y2[i], y2[i], var2);
    }
    f2_unpack(var2, s0, s1);
    const float rstd = 1.0f / sqrtf((s0 + s1) * (1.0f / D) + 1e-5f);
    const uint64_t rstd2 = f2_pack(rstd, rstd);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const ulonglong2 ww = *reinterpret_cast<const ulonglong2 *>(w + k * 4);
        const ulonglong2 bb = *reinterpret_cast<const ulonglong2 *>(b + k * 4);
        y2[2 * k] = f2_fma(y2[2 * k], f2_mul(ww.x, rstd2), bb.x);
        y2[2 * k + 1] = f2_fma(y2[2 * k + 1], f2_mul(ww.y, rstd2), bb.y);
        *reinterpret_cast<ulonglong2 *>(row + k * 4) = make_ulonglong2(y2[2 * k], y2[2 * k + 1]);
    }
}

// the 72 (+8 padding) accumulator columns of my TMEM lane as 36 packed pairs
__device__ __forceinline__ void load_acc_row(uint32_t taddr, uint64_t (&y2)[36]) {
    uint32_t v0[32], v1[32], u[8];
    tmem_ld32(taddr, v0);
    tmem_ld32(taddr + 32, v1);
    tmem_ld8(taddr + 64, u);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        y2[i] = f2_pack(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]));
        y2[16 + i] = f2_pack(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) y2[32 + i] = f2_pack(__uint_as_float(u[2 * i]), __uint_as_float(u[2 * i + 1]));
}

// ---- the fused FFN kernel ---------------------------------------------------------------------------------------------------
// Per tile t (128 tokens) and hidden chunk c the chain is  G1(t,c) -> epilogue(t,c) -> G2(t,c) -> G1(t,c+1) ...; the two tiles'
// chains are issued by two independent warps, so the tensor pipe works on one tile while the other tile's epilogue runs.
//
// OUTPROJ = true prepends the attention output projection of the same encoder layer:  h1 = LN1(h + att · Wo^T + bo)  (one M=128, N=80,
// K=80 MMA block per tile into the Y columns, LayerNorm1 in the epilogue warps), h1 is stored to global (it is LN2's residual) and, as
// fp16, becomes the GEMM1 operand tile in shared memory — so the whole token-wise half of the layer is ONE kernel.
// ---- one-tile variant: CTA = 128 tokens, TWO CTAs per SM ----------------------------------------------------------------------------
// Same chain per hidden chunk (G1 -> epilogue -> G2), but a CTA owns a single M=128 tile and two CTAs share an SM: while one CTA stages
// its operands, runs LayerNorm1 or LayerNorm2 (tensor pipe idle: ~12 k of the two-tile kernel's 38.8 k cycles), the other CTA's MMAs keep
// the tensor pipe busy.  To fit twice in shared memory and TMEM (96.7 KB, 256 columns) the fp32 row slab only exists outside the main loop
// (it overlays weight-ring stages 1..2) and the LayerNorm2 residual does not need it: the output accumulator Y is INITIALISED with
// h1 + b2 (tcgen05.st) and every GEMM2 accumulates onto it.
namespace fast1 {
using namespace fast;
constexpr int TM1 = 128, STG = 3, THREADS1 = 192;  // warp 0 producer (+ TMEM alloc), warp 1 MMA issuer, warps 2-5 epilogue
constexpr int X1_BYTES = KC8 * TM1 * 16;           // 20480
constexpr int OFF1_W = X1_BYTES;
constexpr int OFF1_WO = OFF1_W + STG * STAGE_BYTES;
constexpr int OFF1_PAR = OFF1_WO + WO_BYTES;
constexpr int OFF1_BAR = OFF1_PAR + 6 * D * 4;
constexpr int OFF1_TMEM = OFF1_BAR + 24 * 8;
constexpr int SMEM1 = OFF1_TMEM + 16;
constexpr int SLAB1_BYTES = 4 * 32 * RS * 4;       // 38912: one 32-row slab per epilogue warp, overlaying ring stages 1..2
static_assert(SLAB1_BYTES <= (STG - 1) * STAGE_BYTES, "slab overlays the ring behind stage 0");
static_assert(2 * (SMEM1 + 1024) <= 228 * 1024, "two CTAs per SM");
constexpr int T1_H = 0, T1_Y = 128, T1_COLS = 256;
}  // namespace fast1

template <bool OUTPROJ>
__global__ void __launch_bounds__(fast1::THREADS1, 2)
ffn_ln128_kernel(const float *h_in, float *h_out, const __half *__restrict__ wpack, const float *__restrict__ b2, const float *__restrict__ ln_w,
                 const float *__restrict__ ln_b, int M, int n_chunks, const __half *__restrict__ att_img, const __half *__restrict__ wo_img,
                 const float *__restrict__ bo, const float *__restrict__ ln1_w, const float *__restrict__ ln1_b, float *__restrict__ himg_out, int L,
                 long long *__restrict__ tlog) {
    using namespace fast1;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // bring-up instrumentation (FD_FFN_TLOG=<path>): the first epilogue thread logs clock64() at phase boundaries, 16 slots per CTA
    long long *tl = (tlog && tid == 64) ? tlog + (size_t)blockIdx.x * 16 : nullptr;
    int tli = 0;
#define FD_TLOG() do { if (tl && tli < 14) tl[tli++] = clock64(); } while (0)
    FD_TLOG();  // 0: start
    if (tl) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
        tl[14] = (long long)ns;
    }
    const int m0 = blockIdx.x * TM1;
    const uint32_t bar0 = smem_u32(smem + OFF1_BAR);
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (STG + s); };
    auto H_FULL = [&](int b) { return bar0 + 8u * (2 * STG + b); };
    auto H_READY = [&](int b) { return bar0 + 8u * (2 * STG + 2 + b); };
    const uint32_t Y_FULL = bar0 + 8u * (2 * STG + 4), OP_FULL = bar0 + 8u * (2 * STG + 5), X_READY = bar0 + 8u * (2 * STG + 6),
                   WO_FULL = bar0 + 8u * (2 * STG + 7), X_FULL = bar0 + 8u * (2 * STG + 8), SLAB_FREE = bar0 + 8u * (2 * STG + 9);
    static_assert(2 * STG + 10 <= 24, "barrier block");
    float *par = reinterpret_cast<float *>(smem + OFF1_PAR);
    uint4 *Xs = reinterpret_cast<uint4 *>(smem);  // [kc][128 rows] 16-byte k-chunks
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF1_TMEM);
    const uint32_t w_smem = smem_u32(smem + OFF1_W);
    const uint8_t *wsrc = reinterpret_cast<const uint8_t *>(wpack);
    auto fetch = [&](int c) {  // weight chunk c -> ring stage c % STG
        const int s = c % STG;
        mbar_arrive_expect_tx(W_FULL(s), STAGE_BYTES);
        bulk_g2s(w_smem + s * STAGE_BYTES, wsrc + (size_t)c * STAGE_BYTES, STAGE_BYTES, W_FULL(s));
    };

    if (tid == 0) {
        for (int s = 0; s < STG; ++s) {
            mbar_init(W_FULL(s), 1);
            mbar_init(W_EMPTY(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(H_FULL(b), 1);
            mbar_init(H_READY(b), 128);
        }
        mbar_init(Y_FULL, 1);
        mbar_init(OP_FULL, 1);
        mbar_init(X_READY, 128);
        mbar_init(WO_FULL, 1);
        mbar_init(X_FULL, 1);
        mbar_init(SLAB_FREE, 128);
        mbar_fence_init();
        if (OUTPROJ) {
            // my 128 rows of the attention kernel's fp16 operand image (per 256-token tile [9][256 rows][8 halfs]): nine 2 KB bulk copies
            mbar_arrive_expect_tx(X_FULL, (KC8 - 1) * TM1 * 16);
            const uint8_t *src = reinterpret_cast<const uint8_t *>(att_img) + (size_t)(blockIdx.x >> 1) * ATT_TILE_BYTES + (blockIdx.x & 1) * (TM1 * 16);
            for (int kc = 0; kc < KC8 - 1; ++kc) bulk_g2s(smem_u32(smem) + kc * (TM1 * 16), src + (size_t)kc * (256 * 16), TM1 * 16, X_FULL);
            mbar_arrive_expect_tx(WO_FULL, WO_BYTES);
            bulk_g2s(smem_u32(smem + OFF1_WO), wo_img, WO_BYTES, WO_FULL);
        }
        fetch(0);  // stages 1.. double as the row slabs until LayerNorm1 is done
    }
    if (warp == 0) {
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), T1_COLS);
    }
    if (tid < D) {  // per-column parameters of the two LayerNorm epilogues -> shared memory (broadcast reads)
        par[tid] = OUTPROJ ? bo[tid] : 0.f;
        par[D + tid] = OUTPROJ ? ln1_w[tid] : 0.f;
        par[2 * D + tid] = OUTPROJ ? ln1_b[tid] : 0.f;
        par[3 * D + tid] = b2[tid];
        par[4 * D + tid] = ln_w[tid];
        par[5 * D + tid] = ln_b[tid];
    }
    if (!OUTPROJ) {  // token rows -> fp16 operand tile [kc][row][8 halfs]
        constexpr int ITEMS = (KC8 - 1) * TM1;
        constexpr int PER_THREAD = ITEMS / THREADS1;  // 6
        static_assert(ITEMS % THREADS1 == 0, "tile load split");
        float4 v[PER_THREAD][2];
#pragma unroll
        for (int i = 0; i < PER_THREAD; ++i) {
            const int idx = tid + i * THREADS1;
            const int row = idx % TM1, kc = idx / TM1;
            v[i][0] = v[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + row < M) {
                const float4 *src = reinterpret_cast<const float4 *>(h_in + (size_t)(m0 + row) * D + kc * 8);
                v[i][0] = src[0];
                v[i][1] = src[1];
            }
        }
#pragma unroll
        for (int i = 0; i < PER_THREAD; ++i) Xs[tid + i * THREADS1] = pack8_f16(v[i][0], v[i][1]);
    }
    if (tid < TM1) Xs[(KC8 - 1) * TM1 + tid] = make_uint4(0x3C003C00u, 0u, 0u, 0u);  // k = 72, 73: fp16 1.0 (bias multipliers)
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    FD_TLOG();  // 1: prologue staged

    if (warp == 0) {
        // ===== weight producer =====
        if (lane == 0) {
            mbar_wait(SLAB_FREE, 0);
            for (int c = 1; c < STG && c < n_chunks; ++c) fetch(c);
            for (int c = STG; c < n_chunks; ++c) {
                mbar_wait(W_EMPTY(c % STG), ((c / STG) & 1) ^ 1);
                fetch(c);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (warp-uniform loop, the elected lane issues) =====
        const uint32_t leader = elect_one() ? 1u : 0u;
        const uint32_t idesc1 = make_idesc_f16(128, NC), idesc2 = make_idesc_f16(128, NY);
        const uint32_t tH0 = tmem + T1_H, tY = tmem + T1_Y;
        const uint64_t xd0 = make_smem_desc(smem_u32(smem), TM1 * 16, 128);
        const uint64_t w1d0 = make_smem_desc(w_smem, NC * 16, 128), w2d0 = make_smem_desc(w_smem + W1_BYTES, NY * 16, 128);
        auto gemm1 = [&](int c, int s) {  // H[c&1] = [X | 1 1] · [W1c | b1c]^T
            const uint64_t w1d = w1d0 + (uint64_t)(s * (STAGE_BYTES >> 4));
            const uint32_t tH = tH0 + (c & 1) * NC;
#pragma unroll
            for (int ks = 0; ks < KP / 16; ++ks)
                mma_f16_ss_if(leader, tH, xd0 + (uint64_t)(ks * (2 * TM1 * 16 >> 4)), w1d + (uint64_t)(ks * (2 * NC * 16 >> 4)), idesc1, ks > 0);
            mma_commit_if(leader, H_FULL(c & 1));
        };
        if (OUTPROJ) {  // Y = att · Wo^T, then wait until the epilogue warps have turned it into the LN1 output tile and re-initialised Y
            const uint64_t wod = make_smem_desc(smem_u32(smem + OFF1_WO), NY * 16, 128);
            mbar_wait(X_FULL, 0);
            mbar_wait(WO_FULL, 0);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < KP / 16; ++ks)
                mma_f16_ss_if(leader, tY, xd0 + (uint64_t)(ks * (2 * TM1 * 16 >> 4)), wod + (uint64_t)(ks * (2 * NY * 16 >> 4)), idesc2, ks > 0);
            mma_commit_if(leader, OP_FULL);
        }
        mbar_wait(X_READY, 0);
        tc_fence_after();
        mbar_wait(W_FULL(0), 0);
        tc_fence_after();
        gemm1(0, 0);
        int s = 0, ph = 0;
        for (int c = 0; c < n_chunks; ++c) {
            int s1 = s + 1, ph1 = ph;
            if (s1 == STG) {
                s1 = 0;
                ph1 ^= 1;
            }
            if (c + 1 < n_chunks) {
                mbar_wait(W_FULL(s1), ph1);
                tc_fence_after();
                gemm1(c + 1, s1);
            }
            mbar_wait(H_READY(c & 1), (c >> 1) & 1);
            tc_fence_after();
            const uint64_t w2d = w2d0 + (uint64_t)(s * (STAGE_BYTES >> 4));
            const uint32_t tH = tH0 + (c & 1) * NC;
#pragma unroll
            for (int ks = 0; ks < NC / 16; ++ks)  // Y += relu(H) · W2c^T (Y starts as residual + b2)
                mma_f16_ts_if(leader, tY, tH + ks * 8, w2d + (uint64_t)(ks * (2 * NY * 16 >> 4)), idesc2, 1u);
            mma_commit_if(leader, W_EMPTY(s));
            s = s1;
            ph = ph1;
        }
        mma_commit_if(leader, Y_FULL);
    } else {
        // ===== epilogue warps: TMEM lane quarter warp % 4, thread = token row =====
        const int q = warp & 3;
        const uint32_t lane_base = (uint32_t)(32 * q) << 16;
        const uint32_t tH0 = tmem + lane_base + T1_H, tY = tmem + lane_base + T1_Y;
        const int row0 = m0 + 32 * q;
        float *slab = reinterpret_cast<float *>(smem + OFF1_W + STAGE_BYTES) + (size_t)q * 32 * RS;
        float *row = slab + lane * RS;
        {   // residual rows of h -> slab (coalesced), while the out-proj MMAs run
            const float4 *src = reinterpret_cast<const float4 *>(h_in + (size_t)row0 * D);
            float4 v[KC];
#pragma unroll
            for (int i = 0; i < KC; ++i) {
                const int idx = lane + 32 * i;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row0 + idx / KC < M) v[i] = src[idx];
            }
#pragma unroll
            for (int i = 0; i < KC; ++i) {
                const int idx = lane + 32 * i;
                *reinterpret_cast<float4 *>(slab + (idx / KC) * RS + (idx % KC) * 4) = v[i];
            }
        }
        __syncwarp();
        {
            uint64_t y2[36];
            if (OUTPROJ) {
                mbar_wait(OP_FULL, 0);
                tc_fence_after();
                FD_TLOG();  // 2: out-proj accumulator ready
                load_acc_row(tY, y2);
                residual_layernorm_row(y2, row, par, par + D, par + 2 * D);  // y2 = h1 = LN1(h + att Wo^T + bo)
                const int r = 32 * q + lane;
#pragma unroll
                for (int kc = 0; kc < KC8 - 1; ++kc) {  // fp16 h1 row -> GEMM1 operand tile (my own row only)
                    float e[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) f2_unpack(y2[4 * kc + i], e[2 * i], e[2 * i + 1]);
                    Xs[kc * TM1 + r] = pack8_f16(make_float4(e[0], e[1], e[2], e[3]), make_float4(e[4], e[5], e[6], e[7]));
                }
            } else {
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const ulonglong2 rr = *reinterpret_cast<const ulonglong2 *>(row + k * 4);
                    y2[2 * k] = rr.x;
                    y2[2 * k + 1] = rr.y;
                }
            }
            // Y <- residual + b2: every GEMM2 accumulates onto it, so the slab is not needed again until the result is staged out
            uint32_t v0[32], v1[32], u8[8];
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const ulonglong2 bb = *reinterpret_cast<const ulonglong2 *>(par + 3 * D + k * 4);
                float a0, a1, a2, a3;
                f2_unpack(f2_add(y2[2 * k], bb.x), a0, a1);
                f2_unpack(f2_add(y2[2 * k + 1], bb.y), a2, a3);
                uint32_t *dst = k < 8 ? &v0[4 * k] : k < 16 ? &v1[4 * (k - 8)] : &u8[4 * (k - 16)];
                dst[0] = __float_as_uint(a0);
                dst[1] = __float_as_uint(a1);
                dst[2] = __float_as_uint(a2);
                dst[3] = __float_as_uint(a3);
            }
            tmem_st32(tY, v0);
            tmem_st32(tY + 32, v1);
            tmem_st8(tY + 64, u8);
            tmem_st_wait();
        }
        fence_proxy_async_smem();  // operand-tile writes -> visible to the MMA; my slab reads are ordered before the bulk copies that reuse the stages
        tc_fence_before();
        mbar_arrive(X_READY);
        mbar_arrive(SLAB_FREE);
        FD_TLOG();  // 3: LN1 row in the operand tile, Y initialised
        for (int c = 0; c < n_chunks; ++c) {
            const uint32_t tH = tH0 + (c & 1) * NC;
            mbar_wait(H_FULL(c & 1), (c >> 1) & 1);
            tc_fence_after();
            if ((c & 7) == 0) FD_TLOG();  // 4..7: hidden chunk c = 0, 8, 16, 24 ready
            uint32_t v0[32], v1[32], u[32];
            tmem_ld32(tH, v0);
            tmem_ld32(tH + 32, v1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {  // relu + fp16 pack in one instruction per pair (hidden unit 2j in the low half)
                u[j] = pack_f16x2_relu_sat(__uint_as_float(v0[2 * j + 1]), __uint_as_float(v0[2 * j]));
                u[16 + j] = pack_f16x2_relu_sat(__uint_as_float(v1[2 * j + 1]), __uint_as_float(v1[2 * j]));
            }
            tmem_st32(tH, u);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(H_READY(c & 1));
        }
        // final: Y (= h1 + b2 + FFN) -> LayerNorm2 -> slab (the ring is dead) -> global
        FD_TLOG();  // 8: last hidden chunk handed over
        mbar_wait(Y_FULL, 0);
        tc_fence_after();
        FD_TLOG();  // 9: Y complete
        {
            uint64_t y2[36];
            load_acc_row(tY, y2);
            residual_layernorm_row<false>(y2, row, nullptr, par + 4 * D, par + 5 * D);
            if (himg_out != nullptr && row0 + lane < M) {
                // the next layer's attention kernel stages its token tile with one bulk copy: leave this row in that kernel's tf32 operand
                // image as well — per series [kc][256 positions][4 floats]; consecutive lanes write consecutive 16-byte slots
                const int mtok = row0 + lane, bser = mtok / L, pos = mtok - bser * L;
                uint4 *idst = reinterpret_cast<uint4 *>(himg_out) + (size_t)bser * (KC * 256) + pos;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    float o0, o1, o2, o3;
                    f2_unpack(y2[2 * k], o0, o1);
                    f2_unpack(y2[2 * k + 1], o2, o3);
                    idst[k * 256] = make_uint4(tf32_round_bits(o0), tf32_round_bits(o1), tf32_round_bits(o2), tf32_round_bits(o3));
                }
            }
        }
        __syncwarp();
        {
            float4 *dst = reinterpret_cast<float4 *>(h_out + (size_t)row0 * D);
#pragma unroll
            for (int i = 0; i < KC; ++i) {
                const int idx = lane + 32 * i;
                if (row0 + idx / KC < M) dst[idx] = *reinterpret_cast<const float4 *>(slab + (idx / KC) * RS + (idx % KC) * 4);
            }
        }
    }
    FD_TLOG();  // 10: rows stored
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, T1_COLS);
    FD_TLOG();  // 11: end
    if (tl) {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
        tl[15] = (long long)ns;
    }
#undef FD_TLOG
}

// FD_FFN_TLOG=<path>: per-CTA phase timestamps of the LAST launch are dumped at fd_destroy (bring-up aid, off by default)
static long long *g_ffn_tlog = nullptr;
static long long *ffn_tlog(cudaStream_t s) {
    static const char *path = getenv("FD_FFN_TLOG");
    if (!path) return nullptr;
    if (!g_ffn_tlog) cudaMalloc((void **)&g_ffn_tlog, (size_t)4096 * 16 * sizeof(long long));
    cudaMemsetAsync(g_ffn_tlog, 0, (size_t)4096 * 16 * sizeof(long long), s);
    return g_ffn_tlog;
}
int ffn_dump_tlog() {
    const char *path = getenv("FD_FFN_TLOG");
    if (!path || !g_ffn_tlog) return 0;
    std::vector<long long> host((size_t)4096 * 16);
    cudaDeviceSynchronize();
    cudaMemcpy(host.data(), g_ffn_tlog, host.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    FILE *f = fopen(path, "w");
    if (!f) return 0;
    for (int c = 0; c < 4096; ++c) {
        if (!host[(size_t)c * 16]) continue;
        fprintf(f, "%d", c);
        for (int i = 0; i < 16; ++i) fprintf(f, " %lld", host[(size_t)c * 16 + i]);
        fprintf(f, "\n");
    }
    fclose(f);
    return 0;
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
int fast_path_supported(const fd_config &c) {
    return c.model_kind == FD_MODEL_TRANSFORMER && c.d_model == fast::D && c.d_ff % fast::NC == 0 && c.d_ff >= fast::NC && c.num_layers > 0;
}

int fast_finalize(fd_handle *h) {
    using namespace fast;
    const int ff = h->cfg.d_ff;
    const size_t per_layer = (size_t)(ff / NC) * STAGE_BYTES;
    for (auto &w : h->tl) {
        void *buf = nullptr, *wo = nullptr;
        FD_CUDA(cudaMalloc(&buf, per_layer));
        h->owned.push_back((float *)buf);
        FD_CUDA(cudaMalloc(&wo, WO_BYTES));
        h->owned.push_back((float *)wo);
        pack_ffn_weights_kernel<<<256, 256>>>(w.l1_w, w.l1_b, w.l2_w, (__half *)buf, ff);
        pack_outproj16_kernel<<<16, 256>>>(w.out_w, (__half *)wo);
        FD_CUDA(cudaGetLastError());
        w.l1_pack = (const float *)buf;
        w.out_pack16 = (const float *)wo;
    }
    FD_CUDA(cudaDeviceSynchronize());
    FD_CUDA(cudaFuncSetAttribute(ffn_ln128_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast1::SMEM1));
    FD_CUDA(cudaFuncSetAttribute(ffn_ln128_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast1::SMEM1));
    return 0;
}

// h <- LN2(h + FFN(h)) for layer `layer`, in place on (M, 72) row-major tokens
int launch_ffn_fast(fd_handle *h, int layer, float *hbuf, int M, cudaStream_t s) {
    using namespace fast;
    const TransformerLayerW &w = h->tl[layer];
    ffn_ln128_kernel<false><<<(M + 127) / 128, fast1::THREADS1, fast1::SMEM1, s>>>(hbuf, hbuf, (const __half *)w.l1_pack, w.l2_b, w.n2_w, w.n2_b, M,
                                                                                  h->cfg.d_ff / NC, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                                  h->cfg.max_len, ffn_tlog(s));
    cudaError_t e1 = cudaGetLastError();
    FD_CHECK(e1 == cudaSuccess, "ffn_ln128_kernel launch failed: %s", cudaGetErrorString(e1));
    h->launches += 1;
    g_global_launches += 1;
    return 0;
}

// h <- LN2(h1 + FFN(h1)) with h1 = LN1(h + out_proj(att)): the whole token-wise half of encoder layer `layer` in one kernel
int launch_outproj_ffn_fast(fd_handle *h, int layer, const void *att_img, float *hbuf, int M, float *himg_out, cudaStream_t s) {
    using namespace fast;
    const TransformerLayerW &w = h->tl[layer];
    FD_CHECK(w.out_pack16 != nullptr && att_img != nullptr, "launch_outproj_ffn_fast: out_proj / attention image missing");
    ffn_ln128_kernel<true><<<(M + 127) / 128, fast1::THREADS1, fast1::SMEM1, s>>>(hbuf, hbuf, (const __half *)w.l1_pack, w.l2_b, w.n2_w, w.n2_b, M,
                                                                                 h->cfg.d_ff / NC, (const __half *)att_img, (const __half *)w.out_pack16,
                                                                                 w.out_b, w.n1_w, w.n1_b, himg_out, h->cfg.max_len, ffn_tlog(s));
    cudaError_t e1 = cudaGetLastError();
    FD_CHECK(e1 == cudaSuccess, "ffn_ln128_kernel<outproj> launch failed: %s", cudaGetErrorString(e1));
    h->launches += 1;
    g_global_launches += 1;
    return 0;
}

}  // namespace fd
