// Tensor-core LSTM stack of LSTMScoreModule.forward (score_models.py:309-310: u <- u + LSTM_i(u) for ten independent single-layer
// nn.LSTM(D, D), zero initial state, gate order i,f,g,o), d_model = 72, FD_MATH_TF32 only (the fp32 kernels of fd_generic.cu stay the
// FD_MATH_FP32 reference).
//
// The recurrence is 240 strictly sequential steps per score evaluation (24 positions x 10 layers); a step for one series is a
// (1 x 144) · (144 x 288) product — far too small for tcgen05's 128-row tiles at the batch sizes of this path (cfg 4: 512 series per GPU).
// So a CTA owns 16 series for the whole stack and the step is ONE warp-level m16n8k16 MMA sweep on fp16 operands (11 significant bits,
// like TF32; h is in [-1, 1], the weights are O(0.1), the residual stream saturates at +-65504; fp32 accumulation):
//     gates[16 series][288] = [x_t | h][16][144] · [W_ih | W_hh]^T
// warp w (of 12) owns gate columns [24 w, 24 w + 24): its B fragments (54 registers per thread: 9 k-tiles x 3 n-tiles x 2) stay in REGISTERS
// for the layer, the A fragments (x_t | h of the 16 series, tf32) come from shared memory, stored in fragment order so that a thread
// fetches its four values of a k-tile with one 128-bit load.  The accumulators (+ both biases, fp32) go to shared memory, then the 16 x 72 (series, unit) gate updates run 3 per thread;
// h is written back tf32-rounded as the next step's A operand, the residual u_t += h_t stays fp32 in shared memory.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "fd_common.cuh"

namespace fd {

namespace lt {
constexpr int D = 72, R = 4 * D, S = 16;   // d_model, gate rows, series per CTA
constexpr int NTW = 3;                     // 8-column n-tiles per warp (registers are per SM sub-partition: 3 warps x 32 x 168 <= 16 K)
constexpr int WARPS = R / (8 * NTW);       // 12
constexpr int THREADS = WARPS * 32;        // 384
constexpr int PAIRS = S * D / THREADS;     // (series, unit) gate updates per thread: 3
static_assert(S * D % THREADS == 0, "gate phase split");
constexpr int KT = 2 * D / 16;             // 9 k-tiles of 16 over [x_t | h]
constexpr int AFR = KT * 32 * 4;            // A operand in FRAGMENT order: [k-tile][lane][a0 a1 a2 a3] 32-bit registers of two halfs each
constexpr int GS = R + 4;                  // gate row stride
}  // namespace lt

struct LstmStackW2 {
    const float *w_ih[16], *w_hh[16], *b_ih[16], *b_hh[16];
};

// gate non-linearities on the MUFU unit (tanh.approx: relative error 2^-11, the same order as the TF32 operands of the gate GEMM)
__device__ __forceinline__ float lt_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lt_sigmoid(float x) { return fmaf(0.5f, lt_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ uint32_t lt_h2(float lo, float hi) {  // two fp16 in one register, saturating
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void lt_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(lt::THREADS, 1) lstm_stack_tc_kernel(float *__restrict__ u, LstmStackW2 W, int n_layers, int B, int L) {
    using namespace lt;
    extern __shared__ __align__(16) float lsm[];
    float *xs = lsm;                           // [L][S][D] fp32 layer input / output sequence of my S series
    float *As = xs + (size_t)L * S * D;        // [KT][32][4] tf32 A operand of the current step (x_t | h of the S series), fragment order
    float *gs = As + AFR;                      // [S][GS]   gate pre-activations
    __half *Ah = reinterpret_cast<__half *>(As);
    // element (series si, column k of [x_t | h]) -> half slot of the m16n8k16 A fragment: register a0 (row gid, cols 2 tig, 2 tig + 1),
    // a1 (gid + 8, same cols), a2 (gid, cols 2 tig + 8, + 9), a3 (gid + 8, cols 2 tig + 8, + 9); lane = 4 gid + tig
    auto a_slot = [](int si, int k) {
        const int kk = k & 15;
        return (((((k >> 4) * 32 + (si & 7) * 4 + ((kk & 7) >> 1)) << 2) + (si >> 3) + 2 * (kk >> 3)) << 1) + (kk & 1);
    };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const int n0 = 8 * NTW * warp;             // my gate columns
    for (int b0 = blockIdx.x * S; b0 < B; b0 += gridDim.x * S) {
        __syncthreads();
        for (int idx = tid; idx < L * S * D; idx += THREADS) {  // (s, t, k) coalesced global reads -> xs[t][s][k]
            const int si = idx / (L * D), tk = idx - si * (L * D), t = tk / D, k = tk - t * D;
            xs[((size_t)t * S + si) * D + k] = (b0 + si < B) ? u[((size_t)(b0 + si) * L) * D + tk] : 0.f;
        }
        for (int layer = 0; layer < n_layers; ++layer) {
            // B fragments of m16n8k16 (col-major B = W^T): b0 = W[n][k = 2 tig, 2 tig + 1], b1 = W[n][k = 2 tig + 8, + 9] with n = gate row gid of
            // the n-tile; column k of [W_ih | W_hh]
            uint32_t wf[KT][NTW][2];
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
                for (int nt = 0; nt < NTW; ++nt) {
                    const int n = n0 + 8 * nt + gid;
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {
                        const int k = 16 * kt + 2 * tig + 8 * hb;  // even, so k and k + 1 are on the same side of the x | h boundary (72 is even)
                        const float *src = k < D ? W.w_ih[layer] + (size_t)n * D + k : W.w_hh[layer] + (size_t)n * D + (k - D);
                        wf[kt][nt][hb] = lt_h2(src[0], src[1]);
                    }
                }
            }
            // biases of my accumulator columns (n = n0 + 8 nt + 2 tig + {0, 1}): b_ih + b_hh
            float bias[NTW][2];
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) {
                const int n = n0 + 8 * nt + 2 * tig;
                bias[nt][0] = W.b_ih[layer][n] + W.b_hh[layer][n];
                bias[nt][1] = W.b_ih[layer][n + 1] + W.b_hh[layer][n + 1];
            }
            float cst[PAIRS];  // cell states of my (series, unit) pairs
#pragma unroll
            for (int i = 0; i < PAIRS; ++i) cst[i] = 0.f;
            __syncthreads();
            for (int idx = tid; idx < S * D; idx += THREADS) {  // h <- 0, A <- x_0
                const int si = idx / D, k = idx - si * D;
                Ah[a_slot(si, D + k)] = __float2half_rn(0.f);
                Ah[a_slot(si, k)] = __float2half_rn(fminf(fmaxf(xs[(size_t)si * D + k], -65504.f), 65504.f));
            }
            __syncthreads();
            for (int t = 0; t < L; ++t) {
                float acc[NTW][4];
#pragma unroll
                for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    const uint4 av = reinterpret_cast<const uint4 *>(As)[kt * 32 + lane];
                    const uint32_t a[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                    for (int nt = 0; nt < NTW; ++nt) lt_mma(acc[nt], a, wf[kt][nt][0], wf[kt][nt][1]);
                }
#pragma unroll
                for (int nt = 0; nt < NTW; ++nt) {  // c0:(gid, 2 tig) c1:(gid, 2 tig + 1) c2:(gid + 8, 2 tig) c3:(gid + 8, 2 tig + 1)
                    const int n = n0 + 8 * nt + 2 * tig;
                    *reinterpret_cast<float2 *>(gs + gid * GS + n) = make_float2(acc[nt][0] + bias[nt][0], acc[nt][1] + bias[nt][1]);
                    *reinterpret_cast<float2 *>(gs + (gid + 8) * GS + n) = make_float2(acc[nt][2] + bias[nt][0], acc[nt][3] + bias[nt][1]);
                }
                __syncthreads();
#pragma unroll
                for (int i = 0; i < PAIRS; ++i) {  // (series, unit) pairs tid + THREADS i of the 16 x 72
                    const int pidx = tid + THREADS * i, si = pidx / D, j = pidx - si * D;
                    const float *g = gs + si * GS + j;
                    const float gi = g[0], gf = g[D], gg = g[2 * D], go = g[3 * D];
                    const float ig = lt_sigmoid(gi), fg = lt_sigmoid(gf), og = lt_sigmoid(go);
                    cst[i] = fg * cst[i] + ig * lt_tanh(gg);
                    const float hv = og * lt_tanh(cst[i]);
                    Ah[a_slot(si, D + j)] = __float2half_rn(hv);
                    float *xo = xs + ((size_t)t * S + si) * D + j;
                    *xo = *xo + hv;  // residual; x_t of this layer is not read again
                    if (t + 1 < L) Ah[a_slot(si, j)] = __float2half_rn(fminf(fmaxf(xo[(size_t)S * D], -65504.f), 65504.f));  // next step's x
                }
                __syncthreads();
            }
        }
        for (int idx = tid; idx < L * S * D; idx += THREADS) {
            const int si = idx / (L * D), tk = idx - si * (L * D), t = tk / D, k = tk - t * D;
            if (b0 + si < B) u[((size_t)(b0 + si) * L) * D + tk] = xs[((size_t)t * S + si) * D + k];
        }
    }
}

static size_t lstm_tc_smem(int L) { return ((size_t)L * lt::S * lt::D + lt::AFR + lt::S * lt::GS) * sizeof(float); }

int lstm_stack_tc_supported(const fd_handle *h) {
    const fd_config &c = h->cfg;
    return c.model_kind == FD_MODEL_LSTM && c.d_model == lt::D && c.num_layers <= 16 && c.math_mode == FD_MATH_TF32 &&
           lstm_tc_smem(c.max_len) <= 200 * 1024;
}

int lstm_tc_finalize(fd_handle *h) {
    (void)h;
    FD_CUDA(cudaFuncSetAttribute(lstm_stack_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return 0;
}

int launch_lstm_stack_tc(fd_handle *h, float *u, int B, cudaStream_t s) {
    const fd_config &c = h->cfg;
    LstmStackW2 W;
    for (int i = 0; i < c.num_layers; ++i) {
        W.w_ih[i] = h->ll[i].w_ih;
        W.w_hh[i] = h->ll[i].w_hh;
        W.b_ih[i] = h->ll[i].b_ih;
        W.b_hh[i] = h->ll[i].b_hh;
    }
    int grid = (B + lt::S - 1) / lt::S;
    if (grid > 148) grid = 148;
    lstm_stack_tc_kernel<<<grid, lt::THREADS, lstm_tc_smem(c.max_len), s>>>(u, W, c.num_layers, B, c.max_len);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, lstm_stack_tc_kernel);
        set_error("lstm_stack_tc_kernel launch failed: %s (regs %d, max threads %d, static smem %zu, dynamic smem %zu of max %d)", cudaGetErrorString(e),
                  fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes, lstm_tc_smem(c.max_len), fa.maxDynamicSharedSizeBytes);
        return 1;
    }
    h->launches += 1;
    g_global_launches += 1;
    return 0;
}

}  // namespace fd
