// LSTM score network of the default math mode (LSTMScoreModule.forward, score_models.py:292-317: embed + time row, u <- u + LSTM_i(u) for
// ten independent single-layer nn.LSTM(D, D) with zero initial state and gate order i,f,g,o, unembed), d_model = 72 — and, around it, the
// WHOLE reverse-diffusion loop of the sampler (sampler.py:83-104, sde.py:129-165 / :215-246) in ONE launch.
//
// The recurrence is 240 strictly sequential steps per score evaluation (24 positions x 10 layers at cfg 4) and a step for one series is a
// (1 x 144) . (144 x 288) product: the path is bound by the LATENCY of a step, not by any throughput.  Series never interact, so a CTA
// owns up to 8 series for the entire sampler run — sample, activations and cell states never leave the SM; nothing is launched per
// diffusion step — and a time step is organised to be as short as possible:
//   * roles swapped against the usual batched form: gates^T [288 x 8 series] = [W_ih | W_hh] [288 x 144] . [x_t | h]^T [144 x 8], so the
//     warp-level m16n8k16 MMA (fp16 operands, 11 significant bits like TF32; fp32 accumulate) spends its 16 rows on GATE rows and its 8
//     columns on series: 162 MMAs per step instead of 324 with series on the 16-row side (tcgen05's 128-row tiles would waste > 90 %);
//   * the gate rows are permuted so that m-tile T holds [i | f | g | o] of units 4T .. 4T+3: the four pre-activations of a (unit, series)
//     pair sit in the accumulators of lanes l and l ^ 16, one shuffle pair away — no shared-memory round trip, no second barrier;
//   * warp T owns m-tile T; its 36 A-fragment registers (weights, pre-packed in fragment order by lstm_pack_kernel) stay in registers for
//     the layer; [x_t | h]^T lives in shared memory in B-fragment order (one 64-bit load per k-tile), double-buffered, ONE barrier per step.
// Embed, unembed, the scheduler update (same operation order as sde_step_kernel: bit-identical given the same score) and the Philox
// noise run in the same kernel between score evaluations on plain CUDA cores (2 % of the time).
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "fd_common.cuh"
#include "fd_philox.cuh"
#include "fd_tc.cuh"

namespace fd {

namespace ls {
constexpr int D = 72, KT = 2 * D / 16, MT = 4 * D / 16;  // 9 k-tiles over [x_t | h], 18 m-tiles of gate rows
constexpr int WARPS = MT, THREADS = WARPS * 32;          // 576
constexpr int NS = 8;                                    // series slots of a CTA (the MMA's n dimension)
constexpr int BIMG = KT * 32 * 2;                        // one [x_t | h]^T image: [k-tile][lane][b0 b1] 32-bit words
}  // namespace ls

struct LstmSamplerArgs {
    const uint4 *wfrag;      // [layer][m-tile][k-tile][lane] A fragments (a0..a3), gate rows permuted
    const float *bias;       // [layer][4 D]  b_ih + b_hh, original gate order
    const float *emb_w, *emb_b, *unemb_w, *unemb_b, *G;
    float *x;                // (B, L, C): the sample, read at the start and written back at the end
    float *score_out;        // non-null: ONE evaluation, the score goes here and x is left alone (fd_score)
    const float *temb;       // [n_steps][D] time-embedding rows; temb_per_series: [B][D], one row per series (fd_score_t, n_steps = 1)
    int temb_per_series;
    const float *coef;       // [n_steps][2] drift coefficient on x, diffusion scalar (step_coefficients)
    const float *noise;      // nullptr: Philox; else injected noise [n_steps][B][L][C]
    int n_layers, B, L, C, n_steps, spc, is_ve;  // spc: series per CTA (<= NS)
    int wsm;                 // 1: the weight fragments of the next layer are prefetched into shared memory by a bulk copy (fits for cfg 4)
    float dt, sqrt_dt;
    uint64_t seed, first_series;
    int dbg;                 // timing probes (fd_set_option "lstm_debug"; results are wrong): 1 = no MMAs, 2 = no gate math, 4 = no residual / next-x traffic, 8 = no per-step barrier
};

// gate non-linearities on the MUFU unit (tanh.approx: relative error 2^-11, the same order as the fp16 operands of the gate GEMM)
__device__ __forceinline__ float ls_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ls_sigmoid(float x) { return fmaf(0.5f, ls_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ uint32_t ls_h2(float lo, float hi) {  // two fp16 in one register, saturating
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ __half ls_h(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }
__device__ __forceinline__ void ls_mma(float (&c)[4], const uint4 &a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// half-word slot of element (k of [x_t | h], series n) in a B-fragment image: b0 = (k = 2 tig, 2 tig + 1; n = gid), b1 = (k + 8, k + 9; n = gid)
__device__ __forceinline__ int ls_bslot(int k, int n) {
    const int kk = k & 15;
    return ((((k >> 4) * 32 + n * 4 + ((kk & 7) >> 1)) * 2 + (kk >> 3)) << 1) + (kk & 1);
}

// A fragments of [W_ih | W_hh] with the gate rows permuted: row r of m-tile T = gate r / 4, unit 4 T + r % 4
__global__ void lstm_pack_kernel(const float *__restrict__ w_ih, const float *__restrict__ w_hh, uint4 *__restrict__ frag) {
    using namespace ls;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (m-tile, k-tile, lane)
    if (idx >= MT * KT * 32) return;
    const int lane = idx & 31, kt = (idx >> 5) % KT, T = idx / (32 * KT);
    const int gid = lane >> 2, tig = lane & 3;
    auto w = [&](int r, int k) {  // element (row r of the tile, column k of [W_ih | W_hh])
        const int row = (r >> 2) * D + 4 * T + (r & 3);
        return k < D ? w_ih[(size_t)row * D + k] : w_hh[(size_t)row * D + (k - D)];
    };
    const int k0 = 16 * kt + 2 * tig;
    uint4 f;
    f.x = ls_h2(w(gid, k0), w(gid, k0 + 1));
    f.y = ls_h2(w(gid + 8, k0), w(gid + 8, k0 + 1));
    f.z = ls_h2(w(gid, k0 + 8), w(gid, k0 + 9));
    f.w = ls_h2(w(gid + 8, k0 + 8), w(gid + 8, k0 + 9));
    frag[idx] = f;
}
__global__ void lstm_bias_kernel(const float *__restrict__ b_ih, const float *__restrict__ b_hh, float *__restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = b_ih[i] + b_hh[i];
}

__global__ void __launch_bounds__(ls::THREADS, 1) lstm_sampler_kernel(const LstmSamplerArgs a) {
    using namespace ls;
    extern __shared__ __align__(16) float lsm[];
    const int L = a.L, C = a.C, LC = L * C;
    float *xs = lsm;                                   // [L][NS][D]  activations of my series (fp32 residual stream)
    float *xsm = xs + (size_t)L * NS * D;              // [NS][L][C]  the sample
    float *sc = xsm + (size_t)NS * LC;                 // [NS][L][C]  the score of the current evaluation
    float *weT = sc + (size_t)NS * LC;                 // [C][D]      embedder weight, transposed
    float *wuT = weT + C * D;                          // [D][C]      unembedder weight, transposed
    uint32_t *bimg = reinterpret_cast<uint32_t *>(wuT + D * C);  // [2][KT][32][2]
    __half *bh = reinterpret_cast<__half *>(bimg);
    uint64_t *wbar = reinterpret_cast<uint64_t *>(bimg + 2 * BIMG);  // mbarrier of the weight prefetch
    const uint4 *wsm = reinterpret_cast<const uint4 *>(wbar + 2);    // [MT][KT][32] A fragments of the layer about to run
    const uint32_t wbar_a = tc::smem_u32(wbar), wsm_a = tc::smem_u32(wsm);
    constexpr uint32_t WBYTES = MT * KT * 32 * 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int unit = 4 * warp + (gid & 3);             // my (unit, series slot) of the gate phase
    const int ser = 2 * tig + (gid >> 2);
    const bool lo = gid < 4;                           // lanes gid < 4 hold i and g (series 2 tig), lanes gid >= 4 hold f and o (series 2 tig + 1)
    uint32_t wphase = 0;
    if (a.wsm && tid == 0) {
        tc::mbar_init(wbar_a, 1);
        tc::mbar_fence_init();
    }
    __syncthreads();
    if (a.wsm && tid == 0) {
        tc::mbar_arrive_expect_tx(wbar_a, WBYTES);
        tc::bulk_g2s(wsm_a, a.wfrag, WBYTES, wbar_a);  // layer 0
    }
    for (int i = tid; i < C * D; i += THREADS) {
        const int c = i / D, d = i - c * D;
        weT[i] = a.emb_w[d * C + c];
        wuT[d * C + c] = a.unemb_w[c * D + d];
    }
    for (int i = tid; i < L * NS * D; i += THREADS) xs[i] = 0.f;  // dead series slots stay finite
    for (long long b0 = (long long)blockIdx.x * a.spc; b0 < a.B; b0 += (long long)gridDim.x * a.spc) {
        const int ns = (int)min((long long)a.spc, a.B - b0);  // live series slots
        __syncthreads();
        for (int i = tid; i < NS * LC; i += THREADS) {
            const int s = i / LC;
            xsm[i] = s < ns ? a.x[(size_t)b0 * LC + i] : 0.f;  // (consecutive series are contiguous)
        }
        __syncthreads();
        for (int step = 0; step < a.n_steps; ++step) {
            // ---- embed: xs[t][s][d] = (x[s][t][:] . We[d][:] + be[d]) + temb[d]   (score_models.py:303,306) ----
            const float *temb = a.temb + (size_t)step * D;
            if (a.dbg & 32) {
            } else if ((C & 3) == 0) {
                // thread = (feature d, series slot s): 8 positions at a time, the weight column in registers, the x rows as 128-bit broadcasts
                // (fewer live series than slots: the spare thread groups take every other block of positions)
                const int nsp = ns <= 1 ? 1 : ns <= 2 ? 2 : ns <= 4 ? 4 : 8, rep = NS / nsp;
                const int d = tid % D, sl = (tid / D) % nsp, part = (tid / D) / nsp;  // THREADS = NS * D
                if (sl < ns) {
                    const float bias_d = a.emb_b[d], temb_d = a.temb_per_series ? a.temb[(size_t)(b0 + sl) * D + d] : temb[d];
                    for (int t0 = 8 * part; t0 < L; t0 += 8 * rep) {
                        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        for (int c = 0; c < C; c += 4) {
                            const float w0 = weT[c * D + d], w1 = weT[(c + 1) * D + d], w2 = weT[(c + 2) * D + d], w3 = weT[(c + 3) * D + d];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                if (t0 + i < L) {
                                    const float4 xv = *reinterpret_cast<const float4 *>(xsm + (size_t)(sl * L + t0 + i) * C + c);
                                    acc[i] = fmaf(xv.x, w0, acc[i]);
                                    acc[i] = fmaf(xv.y, w1, acc[i]);
                                    acc[i] = fmaf(xv.z, w2, acc[i]);
                                    acc[i] = fmaf(xv.w, w3, acc[i]);
                                }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (t0 + i < L) xs[((size_t)(t0 + i) * NS + sl) * D + d] = (acc[i] + bias_d) + temb_d;
                    }
                }
            } else {
                for (int o = tid; o < L * NS * D; o += THREADS) {
                    const int tok = o / D, d = o - tok * D, t = tok / NS, sl = tok - t * NS;
                    if (sl >= ns) continue;
                    const float *xr = xsm + (size_t)(sl * L + t) * C;
                    float acc = 0.f;
                    for (int c = 0; c < C; ++c) acc = fmaf(xr[c], weT[c * D + d], acc);
                    xs[o] = (acc + a.emb_b[d]) + (a.temb_per_series ? a.temb[(size_t)(b0 + sl) * D + d] : temb[d]);
                }
            }
            // ---- the LSTM stack (score_models.py:309-310) ----
            for (int layer = 0; layer < a.n_layers; ++layer) {
                uint4 wf[KT];
                const float *bl = a.bias + (size_t)layer * 4 * D;
                const float b_i = bl[unit], b_f = bl[D + unit], b_g = bl[2 * D + unit], b_o = bl[3 * D + unit];
                if (a.wsm) {  // prefetched while the previous layer (or the step boundary) ran
                    tc::mbar_wait(wbar_a, wphase);
                    wphase ^= 1u;
                    const uint4 *wsrc = wsm + (size_t)(warp * KT) * 32 + lane;
#pragma unroll
                    for (int kt = 0; kt < KT; ++kt) wf[kt] = wsrc[kt * 32];
                } else {
                    const uint4 *wsrc = a.wfrag + ((size_t)(layer * MT + warp) * KT) * 32 + lane;
#pragma unroll
                    for (int kt = 0; kt < KT; ++kt) wf[kt] = __ldg(wsrc + kt * 32);
                }
                float cst = 0.f;
                const int slot_h = ls_bslot(D + unit, ser), slot_x = ls_bslot(unit, ser);
                __syncthreads();  // embed / previous layer complete; everybody holds its fragments
                if (a.wsm && tid == 0) {  // next layer's fragments (layer 0 again after the last one: the next score evaluation)
                    tc::fence_proxy_async_smem();
                    tc::mbar_arrive_expect_tx(wbar_a, WBYTES);
                    tc::bulk_g2s(wsm_a, a.wfrag + (size_t)((layer + 1) % a.n_layers) * (MT * KT * 32), WBYTES, wbar_a);
                }
                for (int i = tid; i < NS * D; i += THREADS) {  // image 0: h <- 0, x <- x_0
                    const int s = i / D, k = i - s * D;
                    bh[ls_bslot(D + k, s)] = __float2half_rn(0.f);
                    bh[ls_bslot(k, s)] = ls_h(xs[(size_t)s * D + k]);
                }
                __syncthreads();
                for (int t = 0; t < ((a.dbg & 256) ? 0 : L); ++t) {
                    const uint2 *bsrc = reinterpret_cast<const uint2 *>(bimg + (t & 1) * BIMG) + lane;
                    float *xo = xs + ((size_t)t * NS + ser) * D + unit;
                    const float x_cur = *xo, x_next = (t + 1 < L) ? xo[NS * D] : 0.f;  // issued ahead of the MMAs: off the critical path
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    if (!(a.dbg & 1)) {
                        if (a.dbg & 16) {  // three independent accumulator chains
                            float acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int kt = 0; kt < KT; kt += 3) {
                                const uint2 b0v = bsrc[kt * 32], b1v = bsrc[(kt + 1) * 32], b2v = bsrc[(kt + 2) * 32];
                                ls_mma(acc, wf[kt], b0v.x, b0v.y);
                                ls_mma(acc1, wf[kt + 1], b1v.x, b1v.y);
                                ls_mma(acc2, wf[kt + 2], b2v.x, b2v.y);
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[i] += acc1[i] + acc2[i];
                        } else {
#pragma unroll
                            for (int kt = 0; kt < KT; ++kt) {
                                const uint2 bv = bsrc[kt * 32];
                                ls_mma(acc, wf[kt], bv.x, bv.y);
                            }
                        }
                    }
                    // c0: (row gid, series 2 tig)  c1: (row gid, series 2 tig + 1)  c2 / c3: row gid + 8.  rows 0-3 i, 4-7 f, 8-11 g, 12-15 o
                    const float r0 = __shfl_xor_sync(0xffffffffu, lo ? acc[1] : acc[0], 16);
                    const float r1 = __shfl_xor_sync(0xffffffffu, lo ? acc[3] : acc[2], 16);
                    const float gi = (lo ? acc[0] : r0) + b_i, gf = (lo ? r0 : acc[1]) + b_f;
                    const float gg = (lo ? acc[2] : r1) + b_g, go = (lo ? r1 : acc[3]) + b_o;
                    float hv;
                    if (a.dbg & 2) {
                        hv = gi + gf + gg + go;
                    } else {
                        cst = ls_sigmoid(gf) * cst + ls_sigmoid(gi) * ls_tanh(gg);
                        hv = ls_sigmoid(go) * ls_tanh(cst);
                    }
                    __half *bn = bh + 2 * ((t + 1) & 1) * BIMG;
                    bn[slot_h] = __float2half_rn(hv);
                    if (!(a.dbg & 4)) {
                        *xo = x_cur + hv;  // residual (the layer's input at position t is not read again)
                        if (t + 1 < L) bn[slot_x] = ls_h(x_next);  // next step's x
                    }
                    if (a.dbg & 8) continue;
                    __syncthreads();
                }
            }
            // ---- unembed (score_models.py:313) + scheduler update (sde.py:129-165 / :215-246; operation order of sde_step_kernel) ----
            if (!(a.dbg & 64)) {   // thread = (channel c, token group): 8 tokens at a time, 8 weights in registers, the activation rows as 128-bit broadcasts
                const int TPC = THREADS / C, c = tid % C, tg = tid / C;
                const int n_live = L * ns;  // live tokens, li = l * ns + s
                if (tg < TPC) {
                    const float bias_c = a.unemb_b[c];
                    for (int q0 = tg; q0 < n_live; q0 += 8 * TPC) {
                        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        const float *row[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int li = min(q0 + i * TPC, n_live - 1), l = li / ns, sl = li - l * ns;
                            row[i] = xs + ((size_t)l * NS + sl) * D;
                        }
                        for (int d0 = 0; d0 < D; d0 += 8) {
                            float w8[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) w8[k] = wuT[(d0 + k) * C + c];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 h0 = *reinterpret_cast<const float4 *>(row[i] + d0), h1 = *reinterpret_cast<const float4 *>(row[i] + d0 + 4);
                                acc[i] = fmaf(h0.x, w8[0], acc[i]);
                                acc[i] = fmaf(h0.y, w8[1], acc[i]);
                                acc[i] = fmaf(h0.z, w8[2], acc[i]);
                                acc[i] = fmaf(h0.w, w8[3], acc[i]);
                                acc[i] = fmaf(h1.x, w8[4], acc[i]);
                                acc[i] = fmaf(h1.y, w8[5], acc[i]);
                                acc[i] = fmaf(h1.z, w8[6], acc[i]);
                                acc[i] = fmaf(h1.w, w8[7], acc[i]);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int li = q0 + i * TPC;
                            if (li < n_live) {
                                const int l = li / ns, sl = li - l * ns;
                                sc[(size_t)(sl * L + l) * C + c] = acc[i] + bias_c;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            const float cx = a.coef ? a.coef[2 * step] : 0.f, d0 = a.coef ? a.coef[2 * step + 1] : 0.f;
            const int groups = (LC + 3) / 4;
            for (int it = tid; it < ((a.dbg & 128) ? 0 : ns * groups); it += THREADS) {
                const int sl = it / groups, g = it - sl * groups;
                float zz[4] = {0.f, 0.f, 0.f, 0.f};
                const size_t gs = (size_t)(b0 + sl) * LC;
                if (!a.score_out && !a.noise) normals4(a.seed, a.first_series + (uint64_t)(b0 + sl), (uint32_t)(step + 1), (uint32_t)g, zz);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = 4 * g + j;
                    if (e >= LC) break;
                    const int l = e / C;
                    const float sv = sc[sl * LC + e];
                    if (a.score_out) {
                        a.score_out[gs + e] = sv;
                        continue;
                    }
                    const float zv = a.noise ? a.noise[(size_t)step * a.B * LC + gs + e] : zz[j];
                    const float xv = xsm[sl * LC + e];
                    const float dg = __fmul_rn(d0, a.G[l]);
                    const float dd = __fmul_rn(dg, dg);
                    const float drift = a.is_ve ? -__fmul_rn(dd, sv) : __fsub_rn(__fmul_rn(cx, xv), __fmul_rn(dd, sv));
                    const float av = __fsub_rn(xv, __fmul_rn(drift, a.dt));
                    xsm[sl * LC + e] = __fadd_rn(av, __fmul_rn(a.sqrt_dt, __fmul_rn(dg, zv)));
                }
            }
            __syncthreads();
        }
        if (!a.score_out)
            for (int i = tid; i < ns * LC; i += THREADS) a.x[(size_t)b0 * LC + i] = xsm[i];
    }
    if (a.wsm && tid == 0) tc::mbar_wait(wbar_a, wphase);  // the last prefetch must land before the CTA retires
}

static size_t lstm_sampler_smem(int L, int C, bool wsm) {
    return ((size_t)L * ls::NS * ls::D + 2 * (size_t)ls::NS * L * C + 2 * (size_t)C * ls::D) * sizeof(float) + 2 * ls::BIMG * sizeof(uint32_t) + 16 +
           (wsm ? (size_t)ls::MT * ls::KT * 32 * 16 : 0);
}
constexpr size_t LSTM_SMEM_MAX = 232448;  // 227 KB: the most dynamic shared memory a CTA can opt in to

int lstm_stack_tc_supported(const fd_handle *h) {
    const fd_config &c = h->cfg;
    return c.model_kind == FD_MODEL_LSTM && c.d_model == ls::D && c.num_layers <= 16 && c.math_mode == FD_MATH_TF32 &&
           lstm_sampler_smem(c.max_len, c.n_channels, false) <= LSTM_SMEM_MAX && c.n_channels <= ls::THREADS;
}

int lstm_tc_finalize(fd_handle *h) {
    const fd_config &c = h->cfg;
    if (!lstm_stack_tc_supported(h)) return 0;
    FD_CUDA(cudaFuncSetAttribute(lstm_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LSTM_SMEM_MAX));
    const size_t per_layer = (size_t)ls::MT * ls::KT * 32;  // uint4 fragments
    uint4 *frag = nullptr;
    float *bias = nullptr;
    FD_CUDA(cudaMalloc((void **)&frag, per_layer * c.num_layers * sizeof(uint4)));
    h->owned.push_back((float *)frag);
    FD_CUDA(cudaMalloc((void **)&bias, (size_t)c.num_layers * 4 * ls::D * sizeof(float)));
    h->owned.push_back(bias);
    for (int i = 0; i < c.num_layers; ++i) {
        lstm_pack_kernel<<<(unsigned)((per_layer + 255) / 256), 256>>>(h->ll[i].w_ih, h->ll[i].w_hh, frag + per_layer * i);
        lstm_bias_kernel<<<(4 * ls::D + 255) / 256, 256>>>(h->ll[i].b_ih, h->ll[i].b_hh, bias + (size_t)i * 4 * ls::D, 4 * ls::D);
    }
    FD_CUDA(cudaDeviceSynchronize());
    h->lstm_wfrag = frag;
    h->lstm_bias = bias;
    return 0;
}

// n_steps reverse-diffusion steps on x (score_out == nullptr), or one score evaluation of x into score_out (n_steps = 1, coef unused)
int launch_lstm_sampler(fd_handle *h, float *x, float *score_out, const float *temb, const float *coef, const float *noise, int B, int n_steps,
                        float dt, float sqrt_dt, uint64_t seed, uint64_t first_series, cudaStream_t s) {
    const fd_config &c = h->cfg;
    FD_CHECK(h->lstm_wfrag && h->lstm_bias, "lstm sampler: weights are not packed");
    LstmSamplerArgs a;
    a.wfrag = reinterpret_cast<const uint4 *>(h->lstm_wfrag);
    a.bias = h->lstm_bias;
    a.emb_w = h->emb_w, a.emb_b = h->emb_b, a.unemb_w = h->unemb_w, a.unemb_b = h->unemb_b, a.G = h->G;
    a.x = x, a.score_out = score_out, a.temb = temb, a.coef = coef, a.noise = noise;
    a.temb_per_series = h->temb_per_series;
    FD_CHECK(!a.temb_per_series || (score_out && n_steps == 1), "lstm sampler: per-series times are for single score evaluations");
    a.n_layers = c.num_layers, a.B = B, a.L = c.max_len, a.C = c.n_channels, a.n_steps = n_steps;
    a.is_ve = c.sched_kind == FD_SCHED_VE;
    a.dt = dt, a.sqrt_dt = sqrt_dt, a.seed = seed, a.first_series = first_series;
    a.dbg = h->lstm_debug;
    // a step costs the same whatever the number of live series slots: spread the batch over all SMs first, then fill the slots
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device);
    int spc = (B + sms - 1) / sms;
    spc = spc < 1 ? 1 : spc > ls::NS ? ls::NS : spc;
    a.spc = spc;
    const int grid = std::min((B + spc - 1) / spc, sms);
    a.wsm = lstm_sampler_smem(c.max_len, c.n_channels, true) <= LSTM_SMEM_MAX;
    lstm_sampler_kernel<<<grid, ls::THREADS, lstm_sampler_smem(c.max_len, c.n_channels, a.wsm != 0), s>>>(a);
    cudaError_t e = cudaGetLastError();
    FD_CHECK(e == cudaSuccess, "lstm_sampler_kernel launch failed: %s", cudaGetErrorString(e));
    h->launches += 1;
    g_global_launches += 1;
    return 0;
}

}  // namespace fd
