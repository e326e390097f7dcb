// Persistent encoder-stack kernel of the tensor-core path (d_model = 72, 12 heads, 32 <= max_len <= 256): ALL encoder layers of one
// score evaluation (nn.TransformerEncoder, score_models.py:57-62,87) in ONE launch.
//
// The per-layer kernels of fd_attn.cu / fd_fast.cu run a layer as two grids (attention, then out-proj + LN1 + FFN + LN2); 20 launches
// per score evaluation, each with a partial last wave (1024 or 512 CTAs on 296 slots) and a drain/fill gap.  Here the same two work
// items become TASKS of one resident grid (two CTAs per SM):
//     ATT(layer, series b, head group g)   q|k|v projection + softmax(QK^T)V of 3 heads  -> fp16 operand image of the FFN task
//     FFN(layer, 128-token tile m)         out-proj + LN1 + FFN + LN2                     -> fp32 rows + tf32 operand image of ATT
// CTAs claim tasks from a global queue (atomic counter over a host-built table in topological order) and synchronise through
// dependency counters in global memory (release/acquire at gpu scope): ATT(b) of layer l+1 waits for the FFN tiles that cover series b
// at layer l, FFN(m) waits for the 4 head groups of every series the tile touches.  Series are independent, so nothing else couples
// tasks; no grid-wide barrier exists.  Because a task is only ever claimed by a running CTA and every dependency sits earlier in the
// queue, the scheme cannot deadlock whatever the number of resident CTAs.  The queue order interleaves the FFN tasks of series b - LAG
// with the ATT tasks of series b, so exponential-bound attention CTAs and tensor-bound FFN CTAs share the SMs.
//
// FFN task (new in this file; the attention task is fd_attn.cu's attention_fused_kernel body): 8 epilogue warps — two threads per
// token row, each owning 36 of the 72 columns (row sums exchanged through shared memory) — so the LayerNorm phases take half as long and
// the kernel fits 320 threads x 2 CTAs/SM in registers; the GEMM1 A operand (LayerNorm1 output, fp16) lives in TENSOR MEMORY
// (tcgen05.mma A-from-TMEM), not shared memory: an SS-mode M=128,N=64 MMA would re-read the 128 x 16 A slice from shared memory for
// every 64-unit chunk (6 KB per 32-cycle MMA = 192 B/clk, above the 128 B/clk shared-memory port).
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "fd_common.cuh"
#include "fd_softmax.cuh"
#include "fd_tc.cuh"

namespace fd {

using namespace tc;

namespace stk {
constexpr int D = 72, KC = 18, H = 12, DH = 6, LP = 256;
constexpr int THREADS = 320;                     // warps 0-7: row / epilogue warps, warp 8: MMA issuer, warp 9: control + bulk-copy producer
constexpr int ROW_WARPS = 8;
constexpr int TMEM_COLS = 256;
// ---- attention task: shared-memory layout of attention_fused_kernel (fd_attn.cu) ----
constexpr int VROWS = 8;
constexpr int IMG_Q = 0, IMG_K = 2 * LP * 4, IMG_V = 4 * LP * 4;
constexpr int IMG_FLOATS = IMG_V + (LP / 4) * VROWS * 4;
constexpr int IMG_BYTES = IMG_FLOATS * 4;
constexpr int HPC = 3, NG = H / HPC, NP_G = 80;
constexpr int WG_BYTES = KC * NP_G * 16;         // 23040
constexpr int XS_BYTES = KC * LP * 16;           // 73728
static_assert(XS_BYTES == HPC * IMG_BYTES, "the head images overlay the token tile's slab");
// The token tile and the in_proj weights travel as FP16 images (11 significant bits like tf32; the rows are LayerNorm outputs / embedded
// samples, far inside the fp16 range, conversions saturate): [10 k-chunks of 8 features][256 positions][8 halfs] — k-chunk 9 (features
// 72..79) is zero — in the first 40 KB of the series' 72 KB slab of `himg`.  Half the bytes of the tf32 image the per-layer kernels use,
// and the projection runs kind::f16: 5 k-steps of 16 instead of 9 of 8 at twice the rate.
constexpr int KC16 = 10;
constexpr int XS16_BYTES = KC16 * LP * 16;       // 40960
constexpr int WG16_BYTES = KC16 * NP_G * 16;     // 12800
constexpr int A_WG = XS_BYTES;
constexpr int A_BG = A_WG + WG_BYTES;            // 96768
constexpr int A_MX = A_BG + NP_G * 4;            // 97088: float mx[2][2][128]
constexpr int A_NRM = A_MX + 2 * 2 * 128 * 4;    // 99136: unsigned nrm[2][HPC]
constexpr float BOUNDED_S2 = 14.0f * 14.0f;
// ---- FFN task ----
constexpr int KP = 80, KC8 = KP / 8, NC = 64, NY = 80, STG = 3, TM = 128;
constexpr int W1_BYTES = KC8 * NC * 16, W2_BYTES = (NC / 8) * NY * 16, STAGE_BYTES = W1_BYTES + W2_BYTES;  // 10240 + 10240
constexpr int WO_BYTES = KC8 * NY * 16;          // 12800
constexpr int ATT_TILE_BYTES = (KC8 - 1) * 256 * 16;  // a 256-token tile of the attention output image [9][256][8 halfs]
constexpr int F_ATT = 0;                         // out-proj A operand [10][128][8 halfs] (k-chunk 9 zero)
constexpr int F_WO = F_ATT + KC8 * TM * 16;      // 20480
constexpr int F_RING = F_WO + WO_BYTES;          // 33280: STG weight stages; stages 1.. double as the fp32 row slab outside the main loop
constexpr int F_SLAB = F_RING + STAGE_BYTES;
constexpr int F_PAR = F_RING + STG * STAGE_BYTES;  // 94720: bo | ln1_w | ln1_b | b2 | ln2_w | ln2_b
constexpr int F_RED = F_PAR + 6 * D * 4;         // 96448: float red[4 passes][2 halves][128 rows]
static_assert(TM * D * 4 <= (STG - 1) * STAGE_BYTES, "row slab overlays the ring behind stage 0");
static_assert(F_RED + 4 * 2 * 128 * 4 <= 102400 && A_NRM + 32 <= 102400, "role areas end below the control block");
constexpr int T_H = 0, T_Y = 128, T_X = 208;     // TMEM: hidden chunk x2 (64 each) | Y accumulator (80) | LN1 output as fp16 A operand (40)
static_assert(T_X + KP / 2 <= TMEM_COLS, "TMEM budget");
// ---- control block (both roles) ----
constexpr int CTL = 102400;                      // +0: mbarrier set 0 | +256: tmem slot, task slot | +512: debug counters | +1024: mbarrier set 1
constexpr int SMEM = CTL + 1536;                 // (+512: debug counters, accumulated in shared memory, flushed at kernel end)
constexpr int DBG_SLOTS = 64;                   // per-CTA debug counters: [0..8) totals, [8..36) ATT phase sums, [36..56) FFN phase sums
static_assert(2 * (SMEM + 1024) <= 228 * 1024, "two CTAs per SM");
}  // namespace stk

struct StackLayer {
    const __half *wg_img;            // in_proj fp16 images per head group [10][80][8 halfs] (attn_finalize)
    const float *bg;                 // gathered in_proj bias
    const __half *wpack, *wo_img;    // FFN chunk images, out_proj image (fast_finalize)
    const float *bo, *ln1_w, *ln1_b, *b2, *ln2_w, *ln2_b;
};

constexpr int STK_MAX_LAYERS = 16;

struct StackArgs {
    StackLayer layers[STK_MAX_LAYERS];  // by value: kernel-parameter (constant bank) reads, no dependent global load per task
    int n_layers;
    float *h;                  // (M + pad, 72) fp32 rows: the residual stream
    float *himg;               // per series (slab of 72 KB) the fp16 token image [10][256][8 halfs] (ATT task A operand)
    __half *att_img;           // per 256-token tile [9][256][8 halfs] (FFN task out-proj A operand)
    const uint32_t *table;     // task queue: bit 31 = FFN, bits 24..30 = layer, bits 0..23 = series * 4 + group | tile
    unsigned n_tasks;
    unsigned *next_task;       // monotonic claim counter; this launch owns [task_base, task_base + n_tasks)
    unsigned task_base;
    unsigned *att_done;        // per 128-token tile: completed ATT tasks touching it (monotonic over launches)
    unsigned *ffn_done;        // per series: completed FFN tiles touching it
    unsigned k_base;           // encoder layers completed by earlier launches since the counters were zeroed
    int B, L, M, n_chunks;
    float qscale;
    int allow_bounded;
    int img_primed;            // layer 0 of this launch finds the series images in `himg` (else it gathers its token tile from `h`)
    int flags;                 // bring-up switches (fd_set_option "stack_flags"): 4 = self-test of the bounded-wait post-mortem path
    long long *dbg;            // optional per-CTA cycle counters (fd_set_option "stack_debug"): [8] per CTA, see fd_debug_stack_stats
};

// ---------------------------------------------------------------------------------------------------------------------------------------
// ATT task: body of attention_fused_kernel (fd_attn.cu) with the barriers recycled per task and the dependency protocol around it
// ---------------------------------------------------------------------------------------------------------------------------------------
// Completion of a task is PUBLISHED (gpu-scope fence + relaxed increments of the dependency counters) by the control thread (warp 9,
// lane 0) one task later — after it has issued the next task's bulk copies and passed the set-up barrier, when that warp has nothing
// else to do: a fence right after the task's last stores costs ~2 k cycles (they have to drain to L2 first) and would stall the whole
// CTA at the next block barrier.  Exception: if the next task's dependency is not met yet, the control thread publishes BEFORE it starts
// waiting (a CTA never blocks while it sits on an unpublished completion — otherwise two CTAs could wait for each other).
struct Pending {
    unsigned *ctr;  // nullptr: nothing to publish
    int lo, hi;
};
__device__ __forceinline__ void publish(const Pending &p) {
    if (p.ctr == nullptr) return;
    __threadfence();
    for (int i = p.lo; i <= p.hi; ++i) asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p.ctr + i), "r"(1u) : "memory");
}

// The next task is claimed one task ahead by the control thread, after the set-up barrier of the running task and the publication of the
// previous one — that warp is idle then, so the three dependent L2 round trips (claim counter, queue entry, a first look at the task's
// dependency counter) cost nothing.  `dep_ok`: the dependency was already met at claim time (counters only grow, and the acquire is
// ordered before every thread's accesses of the next task by the block barriers in between), so set-up can skip the poll.
struct Claim {
    unsigned id, entry;
    bool dep_ok;
};
// dependency of a queue entry: counter address and the value it must have reached
__device__ __forceinline__ void dep_of(const StackArgs &a, unsigned e, const unsigned *&ctr, unsigned &target) {
    const int layer = (int)((e >> 24) & 0x7fu), idx = (int)(e & 0xffffffu);
    const unsigned k = a.k_base + (unsigned)layer;
    if (e >> 31) {  // FFN tile: 4 head groups of every series it touches, this layer
        const int m0 = idx * stk::TM;
        const int s_first = m0 / a.L, s_last = min(m0 + stk::TM - 1, a.M - 1) / a.L;
        ctr = a.att_done + idx;
        target = (k + 1u) * 4u * (unsigned)(s_last - s_first + 1);
    } else {  // ATT (series, group): every FFN tile covering the series, previous layer
        const int b = idx >> 2;
        const int t_first = (b * a.L) >> 7, t_last = ((b + 1) * a.L - 1) >> 7;
        ctr = a.ffn_done + b;
        target = k * (unsigned)(t_last - t_first + 1);
    }
}
// The mbarriers are initialised ONCE per CTA (one set per role) and never recycled: every barrier completes a fixed number of phases per
// task, so a task derives the parity to wait for from the number of tasks of its role this CTA has already run (`n_done`).  Re-initialising
// barriers between tasks (mbarrier.inval + init) was measured to be unsafe as well as slow: arrivals are posted operations, and one that
// lands after the word has been invalidated for the next task raises a hardware exception.
// Arrival counts: ATT 0..10 = W_FULL, PROJ_FULL, IMG_READY (8 warps), S_FULL, O_FULL, O_READ (4 warps), P_READY x4 (4 warps), X_FULL x3 (k-chunk thirds of the token tile);
// FFN 0..16 = W_FULL x3, W_EMPTY x3, H_FULL x2, H_READY x2 (8 warps), Y_FULL, OP_FULL, X_READY (8), WO_FULL, ATT_FULL, RES_FULL, SLAB_FREE (8).
__device__ __forceinline__ void init_role_barriers(uint32_t bar0, bool ffn) {
    const unsigned long long lo = ffn ? 0x1113113311111111ull : 0x1112222211311ull, hi = ffn ? 0x3ull : 0ull;  // 1: 1, 2: 4, 3: 8 arrivals (one per row warp: every lane fences, the warp converges, one lane arrives)
    for (int i = 0; i < 17; ++i) {
        const unsigned code = (unsigned)(((i < 16 ? lo : hi) >> (4 * (i & 15))) & 0xfull);
        if (code) mbar_init(bar0 + 8u * i, code == 1 ? 1u : code == 2 ? 4u : 8u);
    }
}
__device__ __forceinline__ void claim_next(const StackArgs &a, Claim &c) {
    c.id = atomicAdd(a.next_task, 1u) - a.task_base;
    c.entry = 0u;
    c.dep_ok = false;
    if (c.id < a.n_tasks) {
        c.entry = __ldg(a.table + c.id);
        const unsigned *ctr;
        unsigned target;
        dep_of(a, c.entry, ctr, target);
        c.dep_ok = (int)(ld_acquire_gpu(ctr) - target) >= 0;
        if (c.dep_ok) fence_proxy_async_all();  // producers' generic-proxy stores -> my bulk copies (async proxy); ~1 k cycles, off the critical path here
    }
}

// Exact two-pass softmax of a (head, query tile) unit: out of line, so that the common regime (bounded heads) keeps a compact hot loop.
// Row maximum over my key half, exchanged with the thread that owns the other half of my row, rounded to an integer shift.
static __device__ __noinline__ float att_exact_shift(uint32_t trow, int hf, int q, int lane, int L, float *slot) {
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    if (128 * hf < L) max_cols16_pipelined<8>(trow, 128 * hf, 128 * hf, L, m0, m1, m2, m3);
    float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    slot[hf * 128 + 32 * q + lane] = m;
    pair_barrier_sync(q);
    m = fmaxf(m, slot[(hf ^ 1) * 128 + 32 * q + lane]);
    return rintf(fminf(fmaxf(m, -4.0e6f), 4.0e6f));
}
static __device__ __noinline__ void att_exact_exp(uint32_t trow, int col, int L, float shift) {
    if (col + 64 <= L) exp_cols16_pipelined<false, 3, 8, 1, 4, 8, 16>(trow, col, col, col, L, shift);
    else if (col < L) exp_cols16_pipelined<true, 3, 8, 1, 4, 8, 16>(trow, col, col, col, L, shift);
}

template <bool FULL, bool DBG>
__device__ __forceinline__ void att_task(uint8_t *smem, const uint32_t tmem, const uint32_t bar0, const StackArgs &a, const StackLayer &w,
                                         const int b, const int g, const unsigned k, const bool use_img, const int tid, const int warp,
                                         const int lane, const bool dep_ok, Pending &pend, Claim &next, const unsigned n_done) {
    using namespace stk;
    const int L = a.L;
    float *Xs = reinterpret_cast<float *>(smem);
    float *bgs = reinterpret_cast<float *>(smem + A_BG);
    float *mx = reinterpret_cast<float *>(smem + A_MX);
    unsigned *nrm = reinterpret_cast<unsigned *>(smem + A_NRM);
    const uint32_t x_smem = smem_u32(smem), wg_smem = smem_u32(smem + A_WG);
    const uint32_t W_FULL = bar0, PROJ_FULL = bar0 + 8, IMG_READY = bar0 + 16, S_FULL = bar0 + 24, O_FULL = bar0 + 32, O_READ = bar0 + 40,
                   P_READY0 = bar0 + 48, X_FULL = bar0 + 80;  // X_FULL + 8 g: third g of the token tile's k-chunks
    const int NT = (L + 127) / 128;
    const int t_first = (b * L) >> 7, t_last = ((b + 1) * L - 1) >> 7;  // 128-token tiles my series touches
    // phases completed by earlier ATT tasks of this CTA: barriers that fire once per task / once per (head, query tile) sub-task
    const uint32_t p1 = n_done & 1u, pn = (n_done * (unsigned)(HPC * NT)) & 1u;
    bool pub_done = false;
    long long *dsm = reinterpret_cast<long long *>(smem + CTL + 512);
    long long *dp = (DBG && tid == 0) ? dsm + 8 : nullptr;
    const long long dt0 = dp ? clock64() : 0;
    int dpi = 0;
#define FD_MARK() do { if (DBG && dp && dpi < 28) dp[dpi++] += clock64() - dt0; } while (0)

    if (warp == ROW_WARPS + 1 && lane == 0) {
        const unsigned dep_target = k * (unsigned)(t_last - t_first + 1) + ((a.flags & 4) ? 1000000u : 0u);  // (flag 4: self-test of the time-out path)
        long long *dq = DBG ? dsm + 56 : nullptr;
        const long long q0 = DBG ? clock64() : 0;
        if (DBG) dq[0] += clock64() - q0;
        if (DBG) dq[1] += clock64() - q0;
        // my series' rows of the previous layer: every FFN tile that covers the series has finished k times since the counters were zeroed
        // (usually already seen satisfied when the task was claimed)
        const long long w0 = DBG ? clock64() : 0;
        if ((!dep_ok || (a.flags & 4)) && (int)(ld_acquire_gpu(a.ffn_done + b) - dep_target) < 0) {
            publish(pend);
            pub_done = true;
            wait_counter_ge(a.ffn_done + b, dep_target, (int)(0x10000000u | (unsigned)((k - a.k_base) << 20) | (unsigned)(b * 4 + g)));
        }
        if (DBG) dsm[5] += clock64() - w0;
        if (DBG) dq[2] += clock64() - q0;
        if (!dep_ok) fence_proxy_async_all();
        if (DBG) dq[3] += clock64() - q0;
        if (use_img) {
            // the weights first, then the token tile in three pieces (k-steps 0-1, 2-3, 4): the projection MMAs of a piece start as soon as
            // it has landed instead of behind the whole copy
            mbar_arrive_expect_tx(W_FULL, WG16_BYTES);
            bulk_g2s(wg_smem, reinterpret_cast<const uint8_t *>(w.wg_img) + (size_t)g * WG16_BYTES, WG16_BYTES, W_FULL);
            const uint8_t *xsrc = reinterpret_cast<const uint8_t *>(a.himg) + (size_t)b * XS_BYTES;
            for (int p3 = 0; p3 < 3; ++p3) {
                const uint32_t off = (uint32_t)p3 * (4 * LP * 16), bytes = p3 < 2 ? 4 * LP * 16 : 2 * LP * 16;
                mbar_arrive_expect_tx(X_FULL + 8u * p3, bytes);
                bulk_g2s(x_smem + off, xsrc + off, bytes, X_FULL + 8u * p3);
            }
        } else {
            for (int p3 = 0; p3 < 3; ++p3) mbar_arrive(X_FULL + 8u * p3);  // keep the barriers' phase count in step with the tasks that do stage an image
            mbar_arrive_expect_tx(W_FULL, WG16_BYTES);
            bulk_g2s(wg_smem, reinterpret_cast<const uint8_t *>(w.wg_img) + (size_t)g * WG16_BYTES, WG16_BYTES, W_FULL);
        }
        if (DBG) dq[4] += clock64() - q0;
    }
    if (tid < NP_G) bgs[tid] = w.bg[g * NP_G + tid];
    if (tid < 2 * HPC) nrm[tid] = 0u;
    if (DBG && tid == 0) dsm[61] += clock64() - dt0;  // thread 0 reaches the set-up barrier
    if (!use_img) {  // first score evaluation after a plain embed: gather the token rows into the fp16 UMMA image [kc][256][8 halfs]
        const float *src = a.h + (size_t)b * L * D;
        constexpr int ITEMS = KC16 * LP;
        constexpr int PER_THREAD = (ITEMS + THREADS - 1) / THREADS;
#pragma unroll 2
        for (int i = 0; i < PER_THREAD; ++i) {
            const int idx = tid + i * THREADS;
            const int row = idx % LP, kc = idx / LP;
            if (idx < ITEMS) {
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (row < L && kc < D / 8) {
                    v0 = __ldcg(reinterpret_cast<const float4 *>(src + (size_t)row * D + kc * 8));
                    v1 = __ldcg(reinterpret_cast<const float4 *>(src + (size_t)row * D + kc * 8 + 4));
                }
                reinterpret_cast<uint4 *>(Xs)[idx] =
                    make_uint4(pack_f16x2_sat(v0.y, v0.x), pack_f16x2_sat(v0.w, v0.z), pack_f16x2_sat(v1.y, v1.x), pack_f16x2_sat(v1.w, v1.z));
            }
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    FD_MARK();  // 0: task set up (barriers, dependency, bulk copies issued)

    if (warp == ROW_WARPS) {
        // ===== MMA issuer =====
        const uint32_t leader = elect_one() ? 1u : 0u;
        {
            const uint32_t idesc_p = make_idesc_f16(128, NP_G);
            // q|k|v = token tile · Wg^T (kind::f16, K = 80 in five k-steps); operands prepared before the copies are awaited
            constexpr uint32_t XHI = smem_desc_hi(128), XSTEP = 2 * LP * 16 >> 4, WSTEP = 2 * NP_G * 16 >> 4;
            uint32_t w_lo = ((wg_smem >> 4) & 0x3FFFu) | (((NP_G * 16u) >> 4) << 16);
            uint32_t x_lo = ((x_smem >> 4) & 0x3FFFu) | (((LP * 16u) >> 4) << 16);
            uint32_t proj_full = PROJ_FULL, d_p = tmem;
            pin_reg(w_lo);
            pin_reg(x_lo);
            pin_reg(proj_full);
            pin_reg(d_p);
            mbar_wait(W_FULL, p1);
#pragma unroll 1
            for (int p3 = 0; p3 < 3; ++p3) {  // k-steps 2 p3, 2 p3 + 1 (the last piece: k-step 4) of both 128-token tiles
                if (use_img) mbar_wait(X_FULL + 8u * p3, p1);
                tc_fence_after();
                const int nks = p3 < 2 ? 2 : 1;
                for (int t = 0; t < NT; ++t)
                    for (int i = 0; i < nks; ++i) {
                        const uint32_t ks = (uint32_t)(2 * p3 + i);
                        mma_f16_ss_lo_if<XHI, XHI>(leader, d_p + t * NP_G, x_lo + (uint32_t)(t * 128 * 16 >> 4) + ks * XSTEP, w_lo + ks * WSTEP, idesc_p,
                                                   ks > 0 ? 1u : 0u, proj_full, (p3 == 2 && t == NT - 1) ? 1u : 0u);
                    }
            }
        }
        const int NK = ((L + 15) / 16) * 16;
        const uint32_t idesc_s = make_idesc_tf32(128, NK), idesc_o = make_idesc_f16(128, 16);
        const int ksteps = (L + 15) / 16, nq = (L + 63) / 64;
        mbar_wait(IMG_READY, p1);
        tc_fence_after();
        int task = 0;
        for (int j = 0; j < HPC; ++j) {
            const uint32_t base = x_smem + j * IMG_BYTES;
            const uint64_t kd = make_smem_desc(base + IMG_K * 4, LP * 16, 128);
            const uint64_t vd = make_smem_desc(base + IMG_V * 4, VROWS * 16, 0);  // SBO 0: rows 8..15 alias rows 0..7
            for (int t = 0; t < NT; ++t, ++task) {
                if (FULL || ksteps == 16) {  // every 64-key quarter and k-step holds real keys (max_len > 240)
                    // operands of the MMAs a barrier releases are prepared and pinned in registers BEFORE the wait (see the FFN issuer)
                    constexpr uint32_t QK_HI = smem_desc_hi(128), V_HI = smem_desc_hi(0);
                    uint32_t q_lo = (((base + IMG_Q * 4 + t * 128 * 16) >> 4) & 0x3FFFu) | (((LP * 16u) >> 4) << 16);
                    uint32_t k_lo = (uint32_t)kd, s_full = S_FULL, d_s = tmem;
                    pin_reg(q_lo);
                    pin_reg(k_lo);
                    pin_reg(s_full);
                    pin_reg(d_s);
                    if (task > 0) {
                        mbar_wait(O_READ, ((task - 1) & 1) ^ pn);
                        tc_fence_after();
                    }
                    mma_tf32_ss_commit_if<QK_HI, QK_HI>(leader, d_s, q_lo, k_lo, idesc_s, s_full);
#pragma unroll
                    for (int qi = 0; qi < 4; ++qi) {
                        const int qt = (qi & 1) * 2 + (qi >> 1);
                        uint32_t lo[4], ta[4], d_o = tmem + 32, o_full = O_FULL;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            lo[k4] = (uint32_t)vd + (uint32_t)((qt * 4 + k4) * (2 * VROWS * 16 >> 4));
                            ta[k4] = tmem + qt * 64 + k4 * 8;
                            pin_reg(lo[k4]);
                            pin_reg(ta[k4]);
                        }
                        pin_reg(d_o);
                        pin_reg(o_full);
                        mbar_wait(P_READY0 + 8u * qt, (task & 1) ^ pn);
                        tc_fence_after();
                        if (qi < 3) mma_f16_ts_x4_if<V_HI>(leader, d_o, ta, lo, idesc_o, qi > 0 ? 1u : 0u);
                        else mma_f16_ts_x4_commit_if<V_HI>(leader, d_o, ta, lo, idesc_o, 1u, o_full);
                    }
                    continue;
                }
                if (task > 0) {
                    mbar_wait(O_READ, ((task - 1) & 1) ^ pn);
                    tc_fence_after();
                }
                const uint64_t qd = make_smem_desc(base + IMG_Q * 4 + t * 128 * 16, LP * 16, 128);
                mma_tf32_ss_if(leader, tmem, qd, kd, idesc_s, 0);
                mma_commit_if(leader, S_FULL);
                for (int qi = 0; qi < 4; ++qi) {
                    const int qt = (qi & 1) * 2 + (qi >> 1);
                    if (qt >= nq) continue;
                    mbar_wait(P_READY0 + 8u * qt, (task & 1) ^ pn);
                    tc_fence_after();
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const int ks = qt * 4 + k4;
                        if (ks < ksteps)
                            mma_f16_ts_if(leader, tmem + 32, tmem + qt * 64 + k4 * 8, vd + (uint64_t)(ks * (2 * VROWS * 16 >> 4)), idesc_o,
                                          (qi > 0 || k4 > 0) ? 1u : 0u);
                    }
                }
                mma_commit_if(leader, O_FULL);
            }
        }
    } else if (warp < ROW_WARPS) {
        // ===== row warps =====
        const int q = warp & 3, hf = warp >> 2;
        const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
        mbar_wait(PROJ_FULL, p1);
        tc_fence_after();
        FD_MARK();  // 1: projection done
        if (hf < NT) {
            const int t = hf;
            const int pos = t * 128 + 32 * q + lane;
            const bool valid = FULL || pos < L;
#pragma unroll
            for (int j = 0; j < HPC; ++j) {
                uint32_t y[3][8];
                tmem_ld8(trow + t * NP_G + 24 * j, y[0]);
                tmem_ld8(trow + t * NP_G + 24 * j + 8, y[1]);
                tmem_ld8(trow + t * NP_G + 24 * j + 16, y[2]);
                tmem_ld_wait();
                float *img = Xs + j * IMG_FLOATS;
                float qv[8], kv[8], vv[8];
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    qv[d] = (valid && d < DH) ? (__uint_as_float(y[0][d]) + bgs[24 * j + d]) * a.qscale : 0.f;
                    kv[d] = (valid && d < DH) ? __uint_as_float(y[1][d]) + bgs[24 * j + 8 + d] : 0.f;
                    vv[d] = (valid && d < DH) ? __uint_as_float(y[2][d]) + bgs[24 * j + 16 + d] : 0.f;
                }
                vv[6] = valid ? 1.0f : 0.f;
                {
                    float qn = 0.f, kn = 0.f;
#pragma unroll
                    for (int d = 0; d < DH; ++d) {
                        qn = fmaf(qv[d], qv[d], qn);
                        kn = fmaf(kv[d], kv[d], kn);
                    }
                    if (!(qn <= 3.0e38f)) qn = 3.0e38f;
                    if (!(kn <= 3.0e38f)) kn = 3.0e38f;
                    const unsigned qb = __reduce_max_sync(0xffffffffu, __float_as_uint(qn)), kb = __reduce_max_sync(0xffffffffu, __float_as_uint(kn));
                    if (lane == 0) {
                        atomicMax(&nrm[j], qb);
                        atomicMax(&nrm[HPC + j], kb);
                    }
                }
                uint4 *qdst = reinterpret_cast<uint4 *>(img + IMG_Q + pos * 4), *kdst = reinterpret_cast<uint4 *>(img + IMG_K + pos * 4);
                qdst[0] = make_uint4(tf32_round_bits(qv[0]), tf32_round_bits(qv[1]), tf32_round_bits(qv[2]), tf32_round_bits(qv[3]));
                qdst[LP] = make_uint4(tf32_round_bits(qv[4]), tf32_round_bits(qv[5]), 0u, 0u);
                kdst[0] = make_uint4(tf32_round_bits(kv[0]), tf32_round_bits(kv[1]), tf32_round_bits(kv[2]), tf32_round_bits(kv[3]));
                kdst[LP] = make_uint4(tf32_round_bits(kv[4]), tf32_round_bits(kv[5]), 0u, 0u);
                __half *vdst = reinterpret_cast<__half *>(img + IMG_V) + (pos / 8) * (VROWS * 8) + (pos % 8);
#pragma unroll
                for (int d = 0; d < 8; ++d) vdst[d * 8] = f32_to_f16_sat(vv[d]);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive_warp(IMG_READY, lane);
        FD_MARK();  // 2: images built
        int task = 0;
        for (int j = 0; j < HPC; ++j) {
            for (int t = 0; t < NT; ++t, ++task) {
                mbar_wait(S_FULL, (task & 1) ^ pn);
                tc_fence_after();
                FD_MARK();  // 3 + 3 task: S ready
                const bool bounded = a.allow_bounded && __uint_as_float(nrm[j]) * __uint_as_float(nrm[HPC + j]) <= BOUNDED_S2;
                float shift = 0.f;
                if (!bounded) shift = att_exact_shift(trow, hf, q, lane, FULL ? 256 : L, mx + (task & 1) * 256);
#pragma unroll 1
                for (int qq = 0; qq < 2; ++qq) {  // my two 64-key quarters; 16-column sub-chunks with the next TMEM load in flight
                    const int col = 128 * hf + 64 * qq;
                    if (bounded) {
                        // no key masking: padded keys have zero v^T rows (ones-row included), so whatever P they get adds nothing to O or to
                        // the denominator, and columns past the last k-step are never read by the P.V MMAs
                        if (FULL || col < L) exp_cols16_pipelined<false, 7, 16, 3, 4, 8, 16>(trow, col, col, col, L, 0.f);
                    } else {
                        att_exact_exp(trow, col, FULL ? 256 : L, shift);
                    }
                    if (FULL || col < L) {
                        tmem_st_wait();
                        tc_fence_before();
                        mbar_arrive_warp(P_READY0 + 8u * (2 * hf + qq), lane);
                    }
                }
                FD_MARK();  // 4 + 3 task: softmax done
                if (hf == 0) {
                    mbar_wait(O_FULL, (task & 1) ^ pn);
                    tc_fence_after();
                    FD_MARK();  // 5 + 3 task: O ready
                    uint32_t o[8];
                    tmem_ld8(trow + 32, o);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive_warp(O_READ, lane);
                    const int qrow = t * 128 + 32 * q + lane;
                    if (qrow < L) {
                        const float inv = 1.0f / __uint_as_float(o[6]);
                        const size_t mrow = (size_t)b * L + qrow;
                        uint8_t *tile = reinterpret_cast<uint8_t *>(a.att_img) + (mrow >> 8) * (size_t)(9 * 256 * 16) + (mrow & 255) * 16;
                        const int c0 = (g * HPC + j) * DH;
#pragma unroll
                        for (int e = 0; e < 3; ++e) {
                            const int c = c0 + 2 * e;
                            *reinterpret_cast<uint32_t *>(tile + (c >> 3) * (256 * 16) + (c & 7) * 2) =
                                pack_f16x2_sat(__uint_as_float(o[2 * e + 1]) * inv, __uint_as_float(o[2 * e]) * inv);
                        }
                    }
                }
            }
        }
    } else if (lane == 0) {  // control thread: nothing else to do during an ATT task
        if (!pub_done) publish(pend);
        claim_next(a, next);
    }
    FD_MARK();  // 21: my rows done
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    FD_MARK();  // 22: all warps done
    pend.ctr = a.att_done;
    pend.lo = t_first;
    pend.hi = t_last;
#undef FD_MARK
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// FFN task: h <- LN2(h1 + W2 relu(W1 h1 + b1) + b2), h1 = LN1(h + att Wo^T + bo) for one 128-token tile
// ---------------------------------------------------------------------------------------------------------------------------------------
// 36 accumulator columns of my TMEM lane as 18 packed fp32 pairs
__device__ __forceinline__ void load_half_row(uint32_t taddr, uint64_t (&y2)[18]) {
    uint32_t v[32], u[4];
    tmem_ld32(taddr, v);
    tmem_ld4(taddr + 32, u);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) y2[i] = f2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
    y2[16] = f2_pack(__uint_as_float(u[0]), __uint_as_float(u[1]));
    y2[17] = f2_pack(__uint_as_float(u[2]), __uint_as_float(u[3]));
}

// LayerNorm over a 72-column row held by TWO threads (36 columns each, warps w and w + 4, same lane): two-pass mean / variance like
// torch's layer_norm, partial sums exchanged through shared memory (`red`: [2 passes][2 halves][128 rows]) with the 64-thread pair barrier.
__device__ __forceinline__ void half_row_layernorm(uint64_t (&y2)[18], const float *w, const float *bia, float *red, int r, int hf, int q) {
    uint64_t s2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 18; ++i) s2 = f2_add(s2, y2[i]);
    float s0, s1;
    f2_unpack(s2, s0, s1);
    const float s = s0 + s1;
    red[hf * 128 + r] = s;
    pair_barrier_sync(q);
    const float so = red[(hf ^ 1) * 128 + r];
    const float mean = (hf ? so + s : s + so) * (1.0f / stk::D);  // same association in both threads
    const uint64_t mean2 = f2_pack(mean, mean);
    uint64_t var2 = f2_pack(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 18; ++i) {
        y2[i] = f2_sub(y2[i], mean2);
        var2 = f2_fma(y2[i], y2[i], var2);
    }
    f2_unpack(var2, s0, s1);
    const float v = s0 + s1;
    red[256 + hf * 128 + r] = v;
    pair_barrier_sync(q);
    const float vo = red[256 + (hf ^ 1) * 128 + r];
    const float rstd = 1.0f / sqrtf((hf ? vo + v : v + vo) * (1.0f / stk::D) + 1e-5f);
    const uint64_t rstd2 = f2_pack(rstd, rstd);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const ulonglong2 ww = *reinterpret_cast<const ulonglong2 *>(w + k * 4);
        const ulonglong2 bb = *reinterpret_cast<const ulonglong2 *>(bia + k * 4);
        y2[2 * k] = f2_fma(y2[2 * k], f2_mul(ww.x, rstd2), bb.x);
        y2[2 * k + 1] = f2_fma(y2[2 * k + 1], f2_mul(ww.y, rstd2), bb.y);
    }
}

template <bool DBG>
__device__ __forceinline__ void ffn_task(uint8_t *smem, const uint32_t tmem, const uint32_t bar0, const StackArgs &a, const StackLayer &w,
                                         const int m, const unsigned k, const bool write_img, const int tid, const int warp, const int lane,
                                         const bool dep_ok, Pending &pend, Claim &next, const unsigned n_done) {
    using namespace stk;
    const int m0 = m * TM, M = a.M, L = a.L, n_chunks = a.n_chunks;
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (STG + s); };
    auto H_FULL = [&](int b) { return bar0 + 8u * (2 * STG + b); };
    auto H_READY = [&](int b) { return bar0 + 8u * (2 * STG + 2 + b); };
    const uint32_t Y_FULL = bar0 + 8u * (2 * STG + 4), OP_FULL = bar0 + 8u * (2 * STG + 5), X_READY = bar0 + 8u * (2 * STG + 6),
                   WO_FULL = bar0 + 8u * (2 * STG + 7), ATT_FULL = bar0 + 8u * (2 * STG + 8), RES_FULL = bar0 + 8u * (2 * STG + 9),
                   SLAB_FREE = bar0 + 8u * (2 * STG + 10);
    static_assert(2 * STG + 11 <= 32, "barrier block");
    float *par = reinterpret_cast<float *>(smem + F_PAR);
    float *red = reinterpret_cast<float *>(smem + F_RED);
    float *slab = reinterpret_cast<float *>(smem + F_SLAB);
    const uint32_t w_smem = smem_u32(smem + F_RING);
    const uint8_t *wsrc = reinterpret_cast<const uint8_t *>(w.wpack);
    auto fetch = [&](int c) {  // weight chunk c -> ring stage c % STG
        const int s = c % STG;
        mbar_arrive_expect_tx(W_FULL(s), STAGE_BYTES);
        bulk_g2s(w_smem + s * STAGE_BYTES, wsrc + (size_t)c * STAGE_BYTES, STAGE_BYTES, W_FULL(s));
    };
    const int s_first = m0 / L, s_last = min(m0 + TM - 1, M - 1) / L;  // series my tile touches
    // phases completed by earlier FFN tasks of this CTA: once-per-task barriers, ring stage s (chunks c = s mod STG), hidden buffer b (c & 1)
    const uint32_t p1 = n_done & 1u;
    auto pw = [&](int s) { return (n_done * (unsigned)((n_chunks - s + STG - 1) / STG)) & 1u; };
    auto phb = [&](int b) { return (n_done * (unsigned)((n_chunks + 1 - b) / 2)) & 1u; };
    bool pub_done = false;
    long long *dsm = reinterpret_cast<long long *>(smem + CTL + 512);
    long long *dp = (DBG && tid == 0) ? dsm + 36 : nullptr;
    const long long dt0 = dp ? clock64() : 0;
    int dpi = 0;
#define FD_MARK() do { if (DBG && dp && dpi < 20) dp[dpi++] += clock64() - dt0; } while (0)

    static_assert(STG == 3, "init_role_barriers assumes three ring stages");
    if (warp == ROW_WARPS + 1 && lane == 0) {
        const unsigned dep_target = (k + 1u) * 4u * (unsigned)(s_last - s_first + 1);
        // the attention output of this layer for every series the tile touches (4 head groups each); transitively also my own rows of the
        // previous layer (the ATT tasks waited for them)
        const long long w0 = DBG ? clock64() : 0;
        if (!dep_ok && (int)(ld_acquire_gpu(a.att_done + m) - dep_target) < 0) {
            publish(pend);
            pub_done = true;
            wait_counter_ge(a.att_done + m, dep_target, (int)(0x20000000u | (unsigned)((k - a.k_base) << 20) | (unsigned)m));
        }
        if (DBG) dsm[6] += clock64() - w0;
        if (!dep_ok) fence_proxy_async_all();
        mbar_arrive_expect_tx(ATT_FULL, (KC8 - 1) * TM * 16);
        const uint8_t *src = reinterpret_cast<const uint8_t *>(a.att_img) + (size_t)(m >> 1) * ATT_TILE_BYTES + (m & 1) * (TM * 16);
        for (int kc = 0; kc < KC8 - 1; ++kc) bulk_g2s(smem_u32(smem + F_ATT) + kc * (TM * 16), src + (size_t)kc * (256 * 16), TM * 16, ATT_FULL);
        mbar_arrive_expect_tx(WO_FULL, WO_BYTES);
        bulk_g2s(smem_u32(smem + F_WO), w.wo_img, WO_BYTES, WO_FULL);
        mbar_arrive_expect_tx(RES_FULL, TM * D * 4);  // my 128 residual rows are one contiguous block (the buffer is padded past M)
        bulk_g2s(smem_u32(slab), a.h + (size_t)m0 * D, TM * D * 4, RES_FULL);
        fetch(0);
    }
    if (tid < D) {
        par[tid] = w.bo[tid];
        par[D + tid] = w.ln1_w[tid];
        par[2 * D + tid] = w.ln1_b[tid];
        par[3 * D + tid] = w.b2[tid];
        par[4 * D + tid] = w.ln2_w[tid];
        par[5 * D + tid] = w.ln2_b[tid];
    }
    if (tid < TM) reinterpret_cast<uint4 *>(smem + F_ATT)[(KC8 - 1) * TM + tid] = make_uint4(0u, 0u, 0u, 0u);  // k = 72..79 of the out-proj operand
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    FD_MARK();  // 0: task set up

    if (warp == ROW_WARPS + 1) {
        // ===== weight producer (first: the previous task's completion) =====
        if (lane == 0) {
            if (!pub_done) publish(pend);
            claim_next(a, next);
            mbar_wait(SLAB_FREE, p1);
            for (int c = 1; c < STG && c < n_chunks; ++c) fetch(c);
            for (int c = STG; c < n_chunks; ++c) {
                mbar_wait(W_EMPTY(c % STG), (((c / STG) & 1) ^ 1) ^ pw(c % STG));
                fetch(c);
            }
        }
    } else if (warp == ROW_WARPS) {
        // ===== MMA issuer =====
        const uint32_t leader = elect_one() ? 1u : 0u;
        const uint32_t idesc1 = make_idesc_f16(128, NC), idesc2 = make_idesc_f16(128, NY);
        const uint32_t tH0 = tmem + T_H, tY = tmem + T_Y, tX = tmem + T_X;
        const uint64_t ad0 = make_smem_desc(smem_u32(smem + F_ATT), TM * 16, 128);
        const uint64_t w1d0 = make_smem_desc(w_smem, NC * 16, 128), w2d0 = make_smem_desc(w_smem + W1_BYTES, NY * 16, 128);
        // B descriptors as (low, high) words: the low words of the MMAs a barrier releases are computed and pinned in registers BEFORE the
        // wait, so only the register -> uniform-register moves and the MMAs themselves follow it (the issuer's instruction stream is on the
        // critical path of the chunk loop: G1(c+1) / G2(c) were ~50 instructions each behind their barrier)
        const uint32_t w1lo0 = (uint32_t)w1d0, w2lo0 = (uint32_t)w2d0;
        constexpr uint32_t WHI = smem_desc_hi(128);  // high word of both weight descriptors
        constexpr uint32_t K1 = 2 * NC * 16 >> 4, K2 = 2 * NY * 16 >> 4, SSTEP = STAGE_BYTES >> 4;
        struct G1Ops {
            uint32_t lo[KP / 16], ta[KP / 16], tH, hfull;
        };
        auto gemm1_prep = [&](int c, int s, G1Ops &o) {  // H[c&1] = [h1 | 1 1] · [W1c | b1c]^T, A = fp16 pairs in TMEM columns [T_X + 8 ks, + 8)
#pragma unroll
            for (int ks = 0; ks < KP / 16; ++ks) {
                o.lo[ks] = w1lo0 + (uint32_t)s * SSTEP + (uint32_t)ks * K1;
                o.ta[ks] = tX + ks * 8;
                pin_reg(o.lo[ks]);
                pin_reg(o.ta[ks]);
            }
            o.tH = tH0 + (c & 1) * NC;
            o.hfull = H_FULL(c & 1);
            pin_reg(o.tH);
            pin_reg(o.hfull);
        };
        auto gemm1_issue = [&](const G1Ops &o, uint32_t wait_bar, uint32_t wait_par) {
            mbar_wait(wait_bar, wait_par);
            tc_fence_after();
            static_assert(KP / 16 == 5 && NC / 16 == 4, "issue helpers");
            mma_f16_ts_x5_commit_if<WHI>(leader, o.tH, o.ta, o.lo, idesc1, 0u, o.hfull);
        };
        {   // Y = att · Wo^T
            const uint64_t wod = make_smem_desc(smem_u32(smem + F_WO), NY * 16, 128);
            mbar_wait(ATT_FULL, p1);
            mbar_wait(WO_FULL, p1);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < KP / 16; ++ks)
                mma_f16_ss_if(leader, tY, ad0 + (uint64_t)(ks * (2 * TM * 16 >> 4)), wod + (uint64_t)(ks * (2 * NY * 16 >> 4)), idesc2, ks > 0);
            mma_commit_if(leader, OP_FULL);
        }
        mbar_wait(X_READY, p1);  // LN1 output in TMEM, Y re-initialised with h1 + b2
        tc_fence_after();
        // Issue order: G1(0), G1(1), then per chunk c: G2(c), G1(c + 2) — G1(c + 2) overwrites the hidden buffer G2(c) has just read (the tensor
        // pipe executes in order).  The operands of BOTH are prepared before the H_READY(c) wait.
        auto stage_of = [&](int c, int &st, uint32_t &par) {  // ring stage and W_FULL parity of chunk c
            st = c % STG;
            par = (uint32_t)((c / STG) & 1) ^ pw(st);
        };
        {
            G1Ops o;
            gemm1_prep(0, 0, o);
            gemm1_issue(o, W_FULL(0), pw(0));
            if (n_chunks > 1) {
                gemm1_prep(1, 1 % STG, o);
                int st;
                uint32_t par;
                stage_of(1, st, par);
                gemm1_issue(o, W_FULL(st), par);
            }
        }
        int s = 0, s2 = 2 % STG;
        uint32_t ph2 = (uint32_t)((2 / STG) & 1);
        for (int c = 0; c < n_chunks; ++c) {
            uint32_t lo[NC / 16], ta[NC / 16], tYp = tY;
#pragma unroll
            for (int ks = 0; ks < NC / 16; ++ks) {  // Y += relu(H) · W2c^T; units 16 ks .. 16 ks + 15 are the packed columns 32 (ks / 2) + 8 (ks % 2) ..
                lo[ks] = w2lo0 + (uint32_t)s * SSTEP + (uint32_t)ks * K2;
                ta[ks] = tH0 + (c & 1) * NC + 32 * (ks >> 1) + 8 * (ks & 1);
                pin_reg(lo[ks]);
                pin_reg(ta[ks]);
            }
            uint32_t wempty = W_EMPTY(s);
            pin_reg(tYp);
            pin_reg(wempty);
            G1Ops o;
            const bool more = c + 2 < n_chunks;
            if (more) gemm1_prep(c + 2, s2, o);
            mbar_wait(H_READY(c & 1), ((c >> 1) & 1) ^ phb(c & 1));
            tc_fence_after();
            mma_f16_ts_x4_commit_if<WHI>(leader, tYp, ta, lo, idesc2, 1u, wempty);
            if (more) gemm1_issue(o, W_FULL(s2), ph2 ^ pw(s2));
            if (++s == STG) s = 0;
            if (++s2 == STG) {
                s2 = 0;
                ph2 ^= 1u;
            }
        }
        mma_commit_if(leader, Y_FULL);
    } else {
        // ===== epilogue warps: TMEM lane quarter q, column half hf; thread = (token row, half) =====
        const int q = warp & 3, hf = warp >> 2, r = 32 * q + lane, token = m0 + r;
        const uint32_t lane_base = (uint32_t)(32 * q) << 16;
        const uint32_t tYh = tmem + lane_base + T_Y + 36 * hf, tX = tmem + lane_base + T_X, tHh = tmem + lane_base + T_H + 32 * hf;
        float *row = slab + r * D + 36 * hf;
        uint64_t y2[18];
        mbar_wait(RES_FULL, p1);
        FD_MARK();  // 1: residual rows staged
        mbar_wait(OP_FULL, p1);
        tc_fence_after();
        FD_MARK();  // 2: out-proj accumulator ready
        load_half_row(tYh, y2);
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) {  // + residual + out-proj bias
            const ulonglong2 rr = *reinterpret_cast<const ulonglong2 *>(row + kk * 4);
            const ulonglong2 bb = *reinterpret_cast<const ulonglong2 *>(par + 36 * hf + kk * 4);
            y2[2 * kk] = f2_add(y2[2 * kk], f2_add(rr.x, bb.x));
            y2[2 * kk + 1] = f2_add(y2[2 * kk + 1], f2_add(rr.y, bb.y));
        }
        half_row_layernorm(y2, par + D + 36 * hf, par + 2 * D + 36 * hf, red, r, hf, q);  // y2 = h1
        {   // h1 -> fp16 A operand of GEMM1 in TMEM (column j = features 2j, 2j+1; column 36 = the two bias multipliers), Y <- h1 + b2
            uint32_t u[16], u16, u17, v[32], v4[4];
#pragma unroll
            for (int i = 0; i < 18; ++i) {
                float e0, e1;
                f2_unpack(y2[i], e0, e1);
                const uint32_t pk = pack_f16x2_sat(e1, e0);
                if (i < 16) u[i] = pk;
                else if (i == 16) u16 = pk;
                else u17 = pk;
            }
            tmem_st16(tX + 18 * hf, u);
            tmem_st2(tX + 18 * hf + 16, u16, u17);
            if (hf) {
                const uint32_t ones[4] = {0x3C003C00u, 0u, 0u, 0u};
                tmem_st4(tX + 36, ones);
            }
#pragma unroll
            for (int kk = 0; kk < 9; ++kk) {
                const ulonglong2 bb = *reinterpret_cast<const ulonglong2 *>(par + 3 * D + 36 * hf + kk * 4);
                float a0, a1, a2, a3;
                f2_unpack(f2_add(y2[2 * kk], bb.x), a0, a1);
                f2_unpack(f2_add(y2[2 * kk + 1], bb.y), a2, a3);
                uint32_t *dst = kk < 8 ? &v[4 * kk] : &v4[0];
                dst[0] = __float_as_uint(a0);
                dst[1] = __float_as_uint(a1);
                dst[2] = __float_as_uint(a2);
                dst[3] = __float_as_uint(a3);
            }
            tmem_st32(tYh, v);
            tmem_st4(tYh + 32, v4);
            tmem_st_wait();
        }
        fence_proxy_async_smem();  // my slab reads are ordered before the bulk copies that reuse ring stages 1..
        tc_fence_before();
        mbar_arrive_warp(X_READY, lane);
        mbar_arrive_warp(SLAB_FREE, lane);
        FD_MARK();  // 3: LN1 done, operand in TMEM
        for (int c = 0; c < n_chunks; ++c) {
            const uint32_t tH = tHh + (c & 1) * NC;
            mbar_wait(H_FULL(c & 1), ((c >> 1) & 1) ^ phb(c & 1));
            tc_fence_after();
            if ((c & 7) == 0) FD_MARK();  // 4..7: hidden chunk 0, 8, 16, 24 ready
            uint32_t v[32], u[16];
            tmem_ld32(tH, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) u[j] = pack_f16x2_relu_sat(__uint_as_float(v[2 * j + 1]), __uint_as_float(v[2 * j]));
            tmem_st16(tH, u);  // my 32 hidden units, packed, into the first 16 of my own 32 columns
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive_warp(H_READY(c & 1), lane);
        }
        FD_MARK();  // 8: last hidden chunk handed over
        mbar_wait(Y_FULL, p1);
        tc_fence_after();
        FD_MARK();  // 9: Y complete
        load_half_row(tYh, y2);  // = h1 + b2 + FFN
        half_row_layernorm(y2, par + 4 * D + 36 * hf, par + 5 * D + 36 * hf, red + 512, r, hf, q);
        if (write_img && token < M) {
            // the next layer's ATT task stages its token tile with one bulk copy: leave my half row in that image too — per series
            // [kc][256 positions][4 floats]; consecutive lanes write consecutive 16-byte slots
            const int bser = token / L, pos = token - bser * L;
            // fp16 image [k-chunk of 8 features][256 positions][8 halfs] in the series' slab: my 36 features are 4.5 chunks — chunk 4 is shared
            // with the other thread of my row (each writes its 8-byte half), the thread of the upper half also zeroes chunk 9
            uint8_t *ibase = reinterpret_cast<uint8_t *>(a.himg) + (size_t)bser * XS_BYTES + (size_t)pos * 16;
            uint32_t pk[18];
#pragma unroll
            for (int i = 0; i < 18; ++i) {
                float e0, e1;
                f2_unpack(y2[i], e0, e1);
                pk[i] = pack_f16x2_sat(e1, e0);  // features 36 hf + 2 i, + 1
            }
            if (hf == 0) {
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8)
                    *reinterpret_cast<uint4 *>(ibase + (size_t)c8 * (256 * 16)) = make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
                *reinterpret_cast<uint2 *>(ibase + (size_t)4 * (256 * 16)) = make_uint2(pk[16], pk[17]);
            } else {
                *reinterpret_cast<uint2 *>(ibase + (size_t)4 * (256 * 16) + 8) = make_uint2(pk[0], pk[1]);
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8)
                    *reinterpret_cast<uint4 *>(ibase + (size_t)(5 + c8) * (256 * 16)) =
                        make_uint4(pk[2 + 4 * c8], pk[3 + 4 * c8], pk[4 + 4 * c8], pk[5 + 4 * c8]);
                *reinterpret_cast<uint4 *>(ibase + (size_t)9 * (256 * 16)) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) *reinterpret_cast<ulonglong2 *>(row + kk * 4) = make_ulonglong2(y2[2 * kk], y2[2 * kk + 1]);
        FD_MARK();  // 10: LN2 done, image + slab written
        asm volatile("bar.sync 5, 256;" ::: "memory");
        {   // the tile's rows are one contiguous block in global memory: coalesced 128-bit stores
            float4 *dst = reinterpret_cast<float4 *>(a.h + (size_t)m0 * D);
            const float4 *srcs = reinterpret_cast<const float4 *>(slab);
#pragma unroll
            for (int i = 0; i < (TM * KC) / 256; ++i) {
                const int idx = tid + 256 * i;
                if (m0 + idx / KC < M) dst[idx] = srcs[idx];
            }
        }
    }
    FD_MARK();  // 11: rows stored
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    FD_MARK();  // 12: all warps done
    pend.ctr = a.ffn_done;
    pend.lo = s_first;
    pend.hi = s_last;
#undef FD_MARK
}

// ---------------------------------------------------------------------------------------------------------------------------------------
template <bool FULL, bool DBG>
__global__ void __launch_bounds__(stk::THREADS, 2) encoder_stack_kernel(const __grid_constant__ StackArgs a) {
    using namespace stk;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_att = smem_u32(smem + CTL), bar_ffn = smem_u32(smem + CTL + 1024);  // one barrier set per role, initialised once
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + CTL + 256);
    volatile unsigned *task_slot = reinterpret_cast<volatile unsigned *>(smem + CTL + 260);  // [0] task id, [1] its queue entry
    if (warp == ROW_WARPS + 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    if (tid == 0) {
        init_role_barriers(bar_att, false);
        init_role_barriers(bar_ffn, true);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    long long *dsm = reinterpret_cast<long long *>(smem + CTL + 512);
    if (DBG && tid < DBG_SLOTS) dsm[tid] = 0;
    Claim next = {0u, 0u, false};
    const bool ctl = warp == ROW_WARPS + 1 && lane == 0;
    unsigned n_att_done = 0, n_ffn_done = 0;  // tasks of each role this CTA has run: the phase offsets of the role's barriers
    if (ctl) claim_next(a, next);
    const long long c_start = DBG ? clock64() : 0;
    long long c_att = 0, c_ffn = 0;
    int n_att = 0, n_ffn = 0;
    Pending pend = {nullptr, 0, 0};
    for (;;) {
        if (ctl) {
            task_slot[0] = next.id;
            task_slot[1] = next.entry;
        }
        __syncthreads();  // also: every warp is done with the previous task
        const unsigned t = task_slot[0], e = task_slot[1];
        if (t >= a.n_tasks) break;
        const bool dep_ok = next.dep_ok;  // (meaningful in the control thread only)
        const int layer = (int)((e >> 24) & 0x7fu), idx = (int)(e & 0xffffffu);
        const StackLayer &w = a.layers[layer];
        const unsigned k = a.k_base + (unsigned)layer;
        const long long c0 = DBG ? clock64() : 0;
        if (e >> 31) {
            ffn_task<DBG>(smem, tmem, bar_ffn, a, w, idx, k, layer + 1 < a.n_layers, tid, warp, lane, dep_ok, pend, next, n_ffn_done);
            ++n_ffn_done;
            if (DBG) {
                c_ffn += clock64() - c0;
                ++n_ffn;
            }
        } else {
            att_task<FULL, DBG>(smem, tmem, bar_att, a, w, idx >> 2, idx & 3, k, layer > 0 || a.img_primed, tid, warp, lane, dep_ok, pend, next, n_att_done);
            ++n_att_done;
            if (DBG) {
                c_att += clock64() - c0;
                ++n_att;
            }
        }
    }
    if (warp == ROW_WARPS + 1 && lane == 0) publish(pend);
    if (DBG && tid == 0) {  // [0] CTA lifetime [1] ATT tasks [2] ATT cycles [3] FFN tasks [4] FFN cycles [5]/[6] dependency waits (ATT/FFN) [7] SM id
        long long *d = a.dbg + (size_t)blockIdx.x * DBG_SLOTS;
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        dsm[0] = clock64() - c_start;
        dsm[1] = n_att;
        dsm[2] = c_att;
        dsm[3] = n_ffn;
        dsm[4] = c_ffn;
        for (int i = 0; i < DBG_SLOTS; ++i) d[i] += dsm[i];
        d[7] = smid;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == ROW_WARPS + 1) tmem_dealloc(tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------------------
// Task queue of one launch: for slot u = 0, 1, ...: the 4 ATT tasks of (layer u / B, series u % B), then the FFN tiles whose LAST series is
// (u - lag) % B of layer (u - lag) / B.  A topological order for 0 <= lag < B - span (span = series a tile can touch beyond its first).
static int build_stack_table(int B, int L, int layers, int lag, std::vector<uint32_t> &out) {
    const int M = B * L, n_tiles = (M + 127) / 128;
    int span = 0;
    for (int m = 0; m < n_tiles; ++m) span = std::max(span, std::min(m * 128 + 127, M - 1) / L - (m * 128) / L);
    out.clear();
    if (lag == -2) {  // layer-major: every ATT task of a layer, then every FFN tile of it (the order of the per-layer kernels)
        for (int layer = 0; layer < layers; ++layer) {
            for (int t = 0; t < 4 * B; ++t) out.push_back(((uint32_t)layer << 24) | (uint32_t)t);
            for (int m = 0; m < n_tiles; ++m) out.push_back(0x80000000u | ((uint32_t)layer << 24) | (uint32_t)m);
        }
        return (int)out.size();
    }
    lag = std::max(0, std::min(lag, B - 1 - span));
    std::vector<std::vector<int>> tiles_of_last(B);
    for (int m = 0; m < n_tiles; ++m) tiles_of_last[std::min(m * 128 + 127, M - 1) / L].push_back(m);
    for (int u = 0; u < layers * B + lag; ++u) {
        if (u < layers * B) {
            const uint32_t layer = u / B, b = u % B;
            for (uint32_t g = 0; g < 4; ++g) out.push_back((layer << 24) | (b * 4 + g));
        }
        if (u >= lag) {
            const uint32_t layer = (u - lag) / B, b = (u - lag) % B;
            for (int m : tiles_of_last[b]) out.push_back(0x80000000u | (layer << 24) | (uint32_t)m);
        }
    }
    return (int)out.size();
}

}  // namespace fd

using namespace fd;

extern "C" int fd_debug_stack_stats(fd_handle *h, int64_t *out, int32_t cap_ctas) {
    if (!h || !h->stk_dbg || !out) return 0;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    const int n = std::min(cap_ctas, h->stk_grid);
    cudaMemcpy(out, h->stk_dbg, (size_t)n * stk::DBG_SLOTS * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemset(h->stk_dbg, 0, (size_t)h->stk_grid * stk::DBG_SLOTS * sizeof(long long));
    return n;
}

extern "C" int fd_stack_task_table(int32_t batch, int32_t max_len, int32_t num_layers, int32_t lag, uint32_t *out, int32_t cap) {
    if (batch <= 0 || max_len <= 0 || num_layers <= 0 || num_layers > 127 || (int64_t)batch * 4 > 0xffffff) return -1;
    std::vector<uint32_t> t;
    const int n = build_stack_table(batch, max_len, num_layers, lag, t);
    if (out)
        for (int i = 0; i < n && i < cap; ++i) out[i] = t[i];
    return n;
}

namespace fd {

int stack_supported(const fd_handle *h) {
    return h->active_path == 1 && h->attn_fast && !h->attn_stream && h->cfg.num_layers <= STK_MAX_LAYERS && h->stack_enabled;
}

void stack_select_slot(fd_handle *h, int slot) {
    if (slot == h->stk_slot) return;
    fd_handle::StkSlot &cur = h->stk_slots[h->stk_slot];
    cur.table = h->stk_table, cur.counters = h->stk_counters, cur.n_tasks = h->stk_n_tasks, cur.table_batch = h->stk_table_batch;
    cur.table_lag = h->stk_table_lag, cur.claims = h->stk_claims, cur.k = h->stk_k;
    const fd_handle::StkSlot &nx = h->stk_slots[slot];
    h->stk_table = nx.table, h->stk_counters = nx.counters, h->stk_n_tasks = nx.n_tasks, h->stk_table_batch = nx.table_batch;
    h->stk_table_lag = nx.table_lag, h->stk_claims = nx.claims, h->stk_k = nx.k;
    h->stk_slot = slot;
}

static int *g_abort_host = nullptr;  // pinned + mapped, one per process

extern "C" int fd_debug_abort_record(int32_t *out8) {
    if (!g_abort_host || !out8) return 0;
    for (int i = 0; i < 8; ++i) out8[i] = ((volatile int *)g_abort_host)[i];
    return g_abort_host[0] != 0;
}

int stack_finalize(fd_handle *h) {
    using namespace stk;
    if (!g_abort_host) {
        FD_CUDA(cudaHostAlloc((void **)&g_abort_host, 64, cudaHostAllocMapped));
        memset(g_abort_host, 0, 64);
    }
    {
        int *dev_view = nullptr;
        FD_CUDA(cudaHostGetDevicePointer((void **)&dev_view, g_abort_host, 0));
        FD_CUDA(cudaMemcpyToSymbol(tc::fd_abort_rec, &dev_view, sizeof(dev_view)));
    }
    FD_CUDA(cudaFuncSetAttribute(encoder_stack_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    FD_CUDA(cudaFuncSetAttribute(encoder_stack_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    FD_CUDA(cudaFuncSetAttribute(encoder_stack_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    FD_CUDA(cudaFuncSetAttribute(encoder_stack_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int sms = 0;
    FD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
    // Two CTAs per SM by construction (shared memory, registers and TMEM are budgeted for it; the occupancy calculator reports 1 because it
    // does not assume the 228 KB shared-memory carve-out, the hardware places 2).  Tasks are claimed dynamically, so a CTA that is not
    // resident yet merely joins later.
    h->stk_grid = 2 * sms;
    return 0;
}

// (re)build the task queue and the dependency counters for `B` series
static int stack_prepare(fd_handle *h, int B, cudaStream_t s) {
    const fd_config &c = h->cfg;
    if (h->stk_table && h->stk_table_batch == B && h->stk_table_lag == h->stack_lag) return 0;
    std::vector<uint32_t> t;
    const int lag = h->stack_lag >= 0 ? h->stack_lag : (h->stack_lag == -2 ? -2 : B / 2);
    const int n = build_stack_table(B, c.max_len, c.num_layers, lag, t);
    const int n_tiles = (B * c.max_len + 127) / 128;
    FD_CUDA(cudaStreamSynchronize(s));  // rare (batch size changed): nobody may still be reading the old queue
    if (h->stk_table) cudaFree(h->stk_table);
    if (h->stk_counters) cudaFree(h->stk_counters);
    h->stk_table = nullptr;
    h->stk_counters = nullptr;
    FD_CUDA(cudaMalloc((void **)&h->stk_table, (size_t)n * sizeof(uint32_t)));
    FD_CUDA(cudaMemcpy(h->stk_table, t.data(), (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    const size_t words = 32 + (size_t)n_tiles + (size_t)B;
    FD_CUDA(cudaMalloc((void **)&h->stk_counters, words * sizeof(unsigned)));
    FD_CUDA(cudaMemset(h->stk_counters, 0, words * sizeof(unsigned)));
    h->stk_n_tasks = n;
    h->stk_table_batch = B;
    h->stk_table_lag = h->stack_lag;
    h->stk_claims = 0;
    h->stk_k = 0;
    return 0;
}

// ws_h <- all encoder layers(ws_h) in one launch.  ws_h / ws_himg / ws_attimg as the per-layer kernels use them.
int launch_encoder_stack(fd_handle *h, int B, cudaStream_t s) {
    using namespace stk;
    const fd_config &c = h->cfg;
    FD_TRY(stack_prepare(h, B, s));
    StackArgs a;
    for (int i = 0; i < c.num_layers; ++i) {
        const TransformerLayerW &w = h->tl[i];
        StackLayer &l = a.layers[i];
        l.wg_img = reinterpret_cast<const __half *>(w.in_pack16);
        l.bg = w.in_bias_pack;
        l.wpack = (const __half *)w.l1_pack;
        l.wo_img = (const __half *)w.out_pack16;
        l.bo = w.out_b;
        l.ln1_w = w.n1_w;
        l.ln1_b = w.n1_b;
        l.b2 = w.l2_b;
        l.ln2_w = w.n2_w;
        l.ln2_b = w.n2_b;
    }
    a.n_layers = c.num_layers;
    a.h = h->ws_h;
    a.himg = h->ws_himg;
    a.att_img = (__half *)h->ws_attimg;
    a.table = h->stk_table;
    a.n_tasks = (unsigned)h->stk_n_tasks;
    a.next_task = h->stk_counters;
    a.task_base = h->stk_claims;
    const int n_tiles = (B * c.max_len + 127) / 128;
    a.att_done = h->stk_counters + 32;
    a.ffn_done = h->stk_counters + 32 + n_tiles;
    a.k_base = h->stk_k;
    a.B = B;
    a.L = c.max_len;
    a.M = B * c.max_len;
    a.n_chunks = c.d_ff / NC;
    a.qscale = (float)(1.4426950408889634 / sqrt((double)DH));
    a.allow_bounded = h->attn_bounded;
    a.img_primed = h->himg_primed && h->himg_fp16;
    a.dbg = nullptr;
    a.flags = h->stack_flags;
    if (h->stack_debug) {
        if (!h->stk_dbg) {
            FD_CUDA(cudaMalloc((void **)&h->stk_dbg, (size_t)h->stk_grid * stk::DBG_SLOTS * sizeof(long long)));
            FD_CUDA(cudaMemset(h->stk_dbg, 0, (size_t)h->stk_grid * stk::DBG_SLOTS * sizeof(long long)));
        }
        a.dbg = h->stk_dbg;
    }
    const int grid = std::min(h->stk_grid, h->stk_n_tasks);
    const bool full = c.max_len == LP;
    if (a.dbg) {
        if (full) encoder_stack_kernel<true, true><<<grid, THREADS, SMEM, s>>>(a);
        else encoder_stack_kernel<false, true><<<grid, THREADS, SMEM, s>>>(a);
    } else {
        if (full) encoder_stack_kernel<true, false><<<grid, THREADS, SMEM, s>>>(a);
        else encoder_stack_kernel<false, false><<<grid, THREADS, SMEM, s>>>(a);
    }
    cudaError_t e = cudaGetLastError();
    FD_CHECK(e == cudaSuccess, "encoder_stack_kernel launch failed: %s", cudaGetErrorString(e));
    h->stk_claims += (unsigned)h->stk_n_tasks + (unsigned)grid;  // every CTA ends on exactly one claim past the queue
    h->stk_k += (unsigned)c.num_layers;
    h->launches += 1;
    g_global_launches += 1;
    return 0;
}

}  // namespace fd
