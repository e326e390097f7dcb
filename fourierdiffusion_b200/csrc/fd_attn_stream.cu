// Tensor-core path, multi-head self-attention for series LONGER than the fused kernel's 256-key tile (fd_attn.cu): d_model = 72, 12 heads
// of 6, any max_len > 256 (US-Droughts' 365, BASELINE cfg 5's 4096, ...).  Same math (nn.TransformerEncoderLayer reached from
// score_models.py:87), two kernels per layer:
//
//   qkv_image_kernel         CTA = (128-token tile, series, 6-head half).  q|k|v = h·Wg^T + bg (tcgen05 kind::tf32, M=128, N=144, K=72);
//                            the epilogue (thread = token) writes UMMA operand IMAGES to global memory — q (pre-scaled by
//                            log2(e)/sqrt(dh)) and k in tf32, v^T in fp16 with a ones-row (column 6 of O becomes the softmax denominator)
//                            — laid out so that the attention kernel stages a 128-query tile (12 KB for its 3 heads) or a 64-key
//                            K|V tile of one head (3 KB) with ONE bulk copy, and reduces max |q|^2, max |k|^2 per (series, head).
//   attention_stream_kernel  CTA = (128-query tile, series, 3-head group), two CTAs per SM.  K|V tiles stream from L2 through an
//                            8-stage bulk-copy ring; S = Q K_tile^T (one MMA, N = 64) rotates through three 64-column TMEM buffers, so
//                            the S of tile u+3 is issued right behind the P·V that consumed buffer u % 3 and the row warps never wait
//                            for the tensor core; P = 2^(S - shift) is written back in place as packed fp16 and O += P·V (kind::f16, A
//                            from TMEM) accumulates in one of two dedicated 16-column accumulators.  Softmax per head: BOUNDED heads
//                            (max|q|·max|k| <= 14 in log2 units) need no row maximum at all; the others make two passes over the keys —
//                            pass 1 recomputes S tile by tile and only tracks the row maximum (Q·K^T has K = 8: the tensor pipe is idle
//                            anyway), pass 2 exponentiates with that maximum — which is exact and needs no O rescaling.
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fd_common.cuh"
#include "fd_softmax.cuh"
#include "fd_tc.cuh"

namespace fd {

using namespace tc;

namespace ats {
constexpr int D = 72, KC = 18, H = 12, DH = 6;
constexpr int HP = 6;                            // heads per projection CTA
constexpr int NPW = HP * 24;                     // projection width: per head q8|k8|v8 (6 real + 2 zero) = 144
constexpr int WP_BYTES = KC * NPW * 16;          // 41472: in_proj image of a 6-head half [kc][144][4]
constexpr int XT_BYTES = KC * 128 * 16;          // 36864: token tile [kc][128][4]
constexpr int HPC = 3, NG = H / HPC;             // heads per attention CTA, head groups
constexpr int Q_FLOATS = 2 * 128 * 4;            // q image of one head and one 128-query tile: [2][128][4]
constexpr int KT = 64;                           // keys per K|V tile
constexpr int K_FLOATS = 2 * KT * 4;             // k image of a tile [2][64][4]
constexpr int V_FLOATS = (KT / 8) * 8 * 8 / 2;   // v^T image of a tile [8][8 rows][8 halfs] = 512 halfs
constexpr int KV_FLOATS = K_FLOATS + V_FLOATS;   // 768 floats = 3072 B
constexpr int KV_BYTES = KV_FLOATS * 4, QG_BYTES = HPC * Q_FLOATS * 4;  // 3072, 12288
constexpr int RST = 8;                           // K|V ring stages
constexpr int ROW_WARPS = 8;
constexpr int THREADS = (ROW_WARPS + 2) * 32;    // + MMA issuer + producer / TMEM allocator
// Measured and dropped (round 2): 48-key tiles with FIVE score buffers and one O accumulator in the same 256 columns (more look-ahead for
// the S of unit u + NBUF behind the P.V of unit u): 170 instead of 150 ns per token at L = 4096, and three 48-key buffers measure the same
// 171 — the look-ahead depth is not what the row warps wait for; the per-unit cost (first TMEM load, store, fence, arrive) is.
constexpr int TMEM_COLS = 256;                   // S buffers [0,64) [64,128) [128,192), O accumulators [192,208) [208,224)
constexpr int OFF_RING = QG_BYTES;
constexpr int OFF_BAR = OFF_RING + RST * KV_BYTES;
constexpr int OFF_TMEM = OFF_BAR + 32 * 8;
constexpr int OFF_MX = OFF_TMEM + 16;            // float mx[2 key halves][128 rows]
constexpr int SMEM_ATT = OFF_MX + 2 * 128 * 4;
constexpr int SMEM_PROJ = XT_BYTES + WP_BYTES + NPW * 4 + 64;
constexpr float BOUNDED_S2 = 14.0f * 14.0f;
}  // namespace ats

// sizes of the global images for `B` series of length L (floats)
static inline size_t ats_tiles(int L) { return (size_t)(L + 127) / 128; }
size_t stream_qimg_floats(int B, int L) { return (size_t)B * ats::NG * ats_tiles(L) * ats::HPC * ats::Q_FLOATS; }
size_t stream_kvimg_floats(int B, int L) { return (size_t)B * ats::NG * ats::HPC * (ats_tiles(L) * 2) * ats::KV_FLOATS; }
size_t stream_nrm_words(int B) { return (size_t)B * ats::H * 2; }

// in_proj weights of the 6-head half hh as the [kc][144][4] UMMA image: column n = (head j = n/24, part p = (n%24)/8 in {q,k,v}, d = n%8);
// real rows of Win are p*72 + (6 hh + j)*6 + d for d < 6, everything else zero.  bias_out[hh][n] is gathered the same way.
__global__ void pack_qkv_half_weights_kernel(const float *__restrict__ w, const float *__restrict__ bias, float *__restrict__ out,
                                             float *__restrict__ bias_out) {
    using namespace ats;
    const int per_half = KC * NPW * 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per_half; i += gridDim.x * blockDim.x) {
        const int hh = i / per_half, e = i % per_half;
        const int j4 = e % 4, n = (e / 4) % NPW, kc = e / (4 * NPW);
        const int j = n / 24, part = (n % 24) / 8, d = n % 8;
        const int row = d < DH ? part * D + (hh * HP + j) * DH + d : -1;
        out[i] = row >= 0 ? __uint_as_float(f32_to_tf32(w[(size_t)row * D + kc * 4 + j4])) : 0.f;
        if (kc == 0 && j4 == 0) bias_out[hh * NPW + n] = row >= 0 ? bias[row] : 0.f;
    }
}

// ---- projection to operand images ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 2)
qkv_image_kernel(const float *__restrict__ h_in, const float *__restrict__ wimg, const float *__restrict__ bimg, float *__restrict__ qimg,
                 float *__restrict__ kvimg, unsigned *__restrict__ nrm, int L, float qscale) {
    using namespace ats;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = blockIdx.x, b = blockIdx.y, hh = blockIdx.z;
    const int NT = gridDim.x, NKT = 2 * NT;
    float *Xs = reinterpret_cast<float *>(smem);
    float *bgs = reinterpret_cast<float *>(smem + XT_BYTES + WP_BYTES);
    const uint32_t x_smem = smem_u32(smem), w_smem = smem_u32(smem + XT_BYTES);
    const uint32_t bar_w = smem_u32(smem + XT_BYTES + WP_BYTES + NPW * 4), bar_mma = bar_w + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + XT_BYTES + WP_BYTES + NPW * 4 + 16);
    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(bar_w, WP_BYTES);
        bulk_g2s(w_smem, reinterpret_cast<const uint8_t *>(wimg) + (size_t)hh * WP_BYTES, WP_BYTES, bar_w);
    }
    if (warp == 4) {
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), 256);
    }
    if (tid < NPW) bgs[tid] = bimg[hh * NPW + tid];
    {   // token rows -> tf32 UMMA image [kc][128][4] (rows >= L zero); a quarter-warp reads 8 consecutive rows of one 16-byte column group
        const float *src = h_in + ((size_t)b * L + (size_t)t * 128) * D;
        for (int idx = tid; idx < KC * 128; idx += 160) {
            const int row = idx % 128, kc = idx / 128;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t * 128 + row < L) v = *reinterpret_cast<const float4 *>(src + (size_t)row * D + kc * 4);
            reinterpret_cast<uint4 *>(Xs)[idx] = make_uint4(tf32_round_bits(v.x), tf32_round_bits(v.y), tf32_round_bits(v.z), tf32_round_bits(v.w));
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 4) {
        mbar_wait(bar_w, 0);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t idesc = make_idesc_tf32(128, NPW);
            const uint64_t a0 = make_smem_desc(x_smem, 128 * 16, 128), b0 = make_smem_desc(w_smem, NPW * 16, 128);
#pragma unroll
            for (int ks = 0; ks < D / 8; ++ks)
                mma_tf32_ss(tmem, a0 + (uint64_t)(ks * (2 * 128 * 16 >> 4)), b0 + (uint64_t)(ks * (2 * NPW * 16 >> 4)), idesc, ks > 0);
            mma_commit(bar_mma);
        }
        __syncwarp();
    } else {
        mbar_wait(bar_mma, 0);
        tc_fence_after();
        const uint32_t trow = tmem + ((uint32_t)(32 * warp) << 16);
        const int r = 32 * warp + lane, pos = t * 128 + r;
        const bool valid = pos < L;
        const int kt = pos / KT, kr = pos % KT;
#pragma unroll 1
        for (int j6 = 0; j6 < HP; ++j6) {
            const int head = hh * HP + j6, g = head / HPC, j = head % HPC;
            uint32_t y[3][8];  // q8 | k8 | v8 of the head
            tmem_ld8(trow + 24 * j6, y[0]);
            tmem_ld8(trow + 24 * j6 + 8, y[1]);
            tmem_ld8(trow + 24 * j6 + 16, y[2]);
            tmem_ld_wait();
            float qv[8], kv[8], vv[8];
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                qv[d] = (valid && d < DH) ? (__uint_as_float(y[0][d]) + bgs[24 * j6 + d]) * qscale : 0.f;
                kv[d] = (valid && d < DH) ? __uint_as_float(y[1][d]) + bgs[24 * j6 + 8 + d] : 0.f;
                vv[d] = (valid && d < DH) ? __uint_as_float(y[2][d]) + bgs[24 * j6 + 16 + d] : 0.f;
            }
            vv[6] = valid ? 1.0f : 0.f;  // ones-row: column 6 of O becomes the softmax denominator
            {
                float qn = 0.f, kn = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) {
                    qn = fmaf(qv[d], qv[d], qn);
                    kn = fmaf(kv[d], kv[d], kn);
                }
                if (!(qn <= 3.0e38f)) qn = 3.0e38f;  // NaN / inf: force the exact path
                if (!(kn <= 3.0e38f)) kn = 3.0e38f;
                const unsigned qb = __reduce_max_sync(0xffffffffu, __float_as_uint(qn)), kb = __reduce_max_sync(0xffffffffu, __float_as_uint(kn));
                if (lane == 0) {
                    atomicMax(&nrm[((size_t)b * H + head) * 2], qb);
                    atomicMax(&nrm[((size_t)b * H + head) * 2 + 1], kb);
                }
            }
            float *qdst = qimg + ((((size_t)b * NG + g) * NT + t) * HPC + j) * Q_FLOATS + r * 4;
            reinterpret_cast<uint4 *>(qdst)[0] = make_uint4(tf32_round_bits(qv[0]), tf32_round_bits(qv[1]), tf32_round_bits(qv[2]), tf32_round_bits(qv[3]));
            reinterpret_cast<uint4 *>(qdst + 128 * 4)[0] = make_uint4(tf32_round_bits(qv[4]), tf32_round_bits(qv[5]), 0u, 0u);
            float *kvt = kvimg + ((((size_t)b * NG + g) * HPC + j) * NKT + kt) * KV_FLOATS;
            reinterpret_cast<uint4 *>(kvt + kr * 4)[0] = make_uint4(tf32_round_bits(kv[0]), tf32_round_bits(kv[1]), tf32_round_bits(kv[2]), tf32_round_bits(kv[3]));
            reinterpret_cast<uint4 *>(kvt + KT * 4 + kr * 4)[0] = make_uint4(tf32_round_bits(kv[4]), tf32_round_bits(kv[5]), 0u, 0u);
            __half *vdst = reinterpret_cast<__half *>(kvt + K_FLOATS) + (kr / 8) * 64 + (kr % 8);  // v^T [key/8][8 rows][8 halfs]
#pragma unroll
            for (int d = 0; d < 8; ++d) vdst[d * 8] = __float2half_rn(fminf(fmaxf(vv[d], -65504.f), 65504.f));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem, 256);
}

// ---- streaming attention ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ats::THREADS, 2)
attention_stream_kernel(const float *__restrict__ qimg, const float *__restrict__ kvimg, const unsigned *__restrict__ nrm, float *__restrict__ att_out,
                        __half *__restrict__ att_img, int L, int allow_bounded) {
    using namespace ats;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = blockIdx.x, b = blockIdx.y, g = blockIdx.z;
    const int NT = gridDim.x, NKT = 2 * NT, nkt = (L + KT - 1) / KT;
    float *mx = reinterpret_cast<float *>(smem + OFF_MX);
    const uint32_t q_smem = smem_u32(smem), ring = smem_u32(smem + OFF_RING);
    const uint32_t bar0 = smem_u32(smem + OFF_BAR);
    const uint32_t Q_FULL = bar0;
    auto KV_FULL = [&](int s) { return bar0 + 8u * (1 + s); };
    auto KV_EMPTY = [&](int s) { return bar0 + 8u * (1 + RST + s); };
    auto S_FULL = [&](int buf) { return bar0 + 8u * (1 + 2 * RST + buf); };
    auto P_READY = [&](int buf) { return bar0 + 8u * (4 + 2 * RST + buf); };
    auto O_FULL = [&](int ob) { return bar0 + 8u * (7 + 2 * RST + ob); };
    auto O_READ = [&](int ob) { return bar0 + 8u * (9 + 2 * RST + ob); };
    static_assert(11 + 2 * RST <= 32, "barrier block");
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_TMEM);
    // per head: bounded (one pass) or exact (two passes over the keys); CTA-uniform, every role derives the same unit sequence from it
    bool bnd[HPC];
    int units_of[HPC], U = 0;
#pragma unroll
    for (int j = 0; j < HPC; ++j) {
        const unsigned *n2 = nrm + ((size_t)b * H + g * HPC + j) * 2;
        bnd[j] = allow_bounded && __uint_as_float(n2[0]) * __uint_as_float(n2[1]) <= BOUNDED_S2;
        units_of[j] = bnd[j] ? nkt : 2 * nkt;
        U += units_of[j];
    }
    // unit u -> (head j, pass-1 flag, key tile)
    auto decode = [&](int u, int &j, bool &maxpass, int &kt) {
        j = 0;
        while (u >= units_of[j]) {
            u -= units_of[j];
            ++j;
        }
        maxpass = !bnd[j] && u < nkt;
        kt = u < nkt ? u : u - nkt;
    };
    const float *kv_base = kvimg + (((size_t)b * NG + g) * HPC) * NKT * KV_FLOATS;

    if (tid == 0) {
        mbar_init(Q_FULL, 1);
        for (int s = 0; s < RST; ++s) {
            mbar_init(KV_FULL(s), 1);
            mbar_init(KV_EMPTY(s), 1);
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init(S_FULL(i), 1);
            mbar_init(P_READY(i), 4);  // one arrival per row warp of the unit's group
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(O_FULL(i), 1);
            mbar_init(O_READ(i), 128);
        }
        mbar_fence_init();
        mbar_arrive_expect_tx(Q_FULL, QG_BYTES);
        bulk_g2s(q_smem, reinterpret_cast<const uint8_t *>(qimg + (((size_t)b * NG + g) * NT + t) * HPC * Q_FLOATS), QG_BYTES, Q_FULL);
    }
    if (warp == ROW_WARPS + 1) {
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == ROW_WARPS + 1) {
        // ===== K|V producer: one 3 KB tile per unit through the ring =====
        if (lane == 0) {
            for (int u = 0; u < U; ++u) {
                int j, kt;
                bool mp;
                decode(u, j, mp, kt);
                const int s = u % RST;
                if (u >= RST) mbar_wait(KV_EMPTY(s), ((u / RST) & 1) ^ 1);
                mbar_arrive_expect_tx(KV_FULL(s), KV_BYTES);
                bulk_g2s(ring + s * KV_BYTES, kv_base + ((size_t)j * NKT + kt) * KV_FLOATS, KV_BYTES, KV_FULL(s));
            }
        }
    } else if (warp == ROW_WARPS) {
        // ===== MMA issuer (warp-uniform, the elected lane issues) =====
        const uint32_t leader = elect_one() ? 1u : 0u;
        const uint32_t idesc_s = make_idesc_tf32(128, KT), idesc_o = make_idesc_f16(128, 16);
        mbar_wait(Q_FULL, 0);
        tc_fence_after();
        // Unit sequence as a cursor (head j, pass-1 flag, key tile) instead of a search per unit; every operand of the MMAs a barrier releases
        // is prepared and pinned in registers BEFORE the wait, one asm block per GEMM: the issuer's instruction stream sits on the critical
        // path of every unit (S of unit u + 3 follows the P.V of unit u), ~50 instructions behind each barrier before.
        struct Cursor {
            int j, kt;
            bool mp;
        };
        auto advance = [&](Cursor &c) {
            if (++c.kt == nkt) {
                c.kt = 0;
                if (c.mp) {
                    c.mp = false;
                } else {
                    ++c.j;
                    c.mp = c.j < HPC && !bnd[c.j];
                }
            }
        };
        constexpr uint32_t QK_HI = smem_desc_hi(128), V_HI = smem_desc_hi(0);
        const uint32_t q_lo0 = ((q_smem >> 4) & 0x3FFFu) | (((128 * 16u) >> 4) << 16);
        const uint32_t k_lo0 = ((ring >> 4) & 0x3FFFu) | (((KT * 16u) >> 4) << 16);
        const uint32_t v_lo0 = (((ring + K_FLOATS * 4) >> 4) & 0x3FFFu) | (((8 * 16u) >> 4) << 16);
        auto issue_s = [&](const Cursor &c, int s, uint32_t kv_par, int buf) {  // S of the cursor's unit into score buffer `buf`, K from ring stage s
            uint32_t q_lo = q_lo0 + (uint32_t)c.j * ((Q_FLOATS * 4) >> 4), k_lo = k_lo0 + (uint32_t)s * (KV_BYTES >> 4);
            uint32_t d_s = tmem + 64 * buf, bar = S_FULL(buf);
            pin_reg(q_lo);
            pin_reg(k_lo);
            pin_reg(d_s);
            pin_reg(bar);
            mbar_wait(KV_FULL(s), kv_par);
            tc_fence_after();
            mma_tf32_ss_commit_if<QK_HI, QK_HI>(leader, d_s, q_lo, k_lo, idesc_s, bar);
        };
        Cursor cs = {0, 0, !bnd[0]};  // unit whose S is issued next
        int ss = 0, sbuf = 0;
        uint32_t spar = 0;
        auto next_s = [&]() {
            issue_s(cs, ss, spar, sbuf);
            advance(cs);
            if (++ss == RST) {
                ss = 0;
                spar ^= 1u;
            }
            if (++sbuf == 3) sbuf = 0;
        };
        for (int u = 0; u < 3 && u < U; ++u) next_s();
        Cursor cu = {0, 0, !bnd[0]};
        int s = 0, buf = 0;
        uint32_t ppar = 0;
        for (int u = 0; u < U; ++u) {
            const int ob = cu.j & 1;
            uint32_t lo[4], ta[4], d_o = tmem + 192 + 16 * ob, kv_empty = KV_EMPTY(s);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                // the tile's two 32-key chunks leave their packed P at columns [0,16) and [32,48) of the buffer
                lo[k4] = v_lo0 + (uint32_t)s * (KV_BYTES >> 4) + (uint32_t)(k4 * (2 * 8 * 16 >> 4));
                ta[k4] = tmem + 64 * buf + (k4 >> 1) * 32 + (k4 & 1) * 8;
                pin_reg(lo[k4]);
                pin_reg(ta[k4]);
            }
            pin_reg(d_o);
            pin_reg(kv_empty);
            if (!cu.mp && cu.kt == 0 && cu.j >= 2) {  // the O accumulator's previous tenant (head j - 2) must have been read out
                mbar_wait(O_READ(ob), 0);
                tc_fence_after();
            }
            mbar_wait(P_READY(buf), ppar);
            tc_fence_after();
            if (!cu.mp) {
                const int nks = min(4, (L - cu.kt * KT + 15) / 16);  // k-steps of 16 keys that hold real keys
                const uint32_t acc = cu.kt > 0 ? 1u : 0u;
                if (nks == 4) {
                    mma_f16_ts_x4_if<V_HI>(leader, d_o, ta, lo, idesc_o, acc);
                } else {
                    const uint64_t vd = make_smem_desc(ring + s * KV_BYTES + K_FLOATS * 4, 8 * 16, 0);  // v^T rows 8..15 alias rows 0..7 (SBO 0)
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4)
                        if (k4 < nks) mma_f16_ts_if(leader, d_o, ta[k4], vd + (uint64_t)(k4 * (2 * 8 * 16 >> 4)), idesc_o, (cu.kt > 0 || k4 > 0) ? 1u : 0u);
                }
                if (cu.kt == nkt - 1) mma_commit_if(leader, O_FULL(ob));
            }
            mma_commit_if(leader, kv_empty);  // arrives once S(u) and P·V(u) have read the stage
            if (u + 3 < U) next_s();
            advance(cu);
            if (++s == RST) s = 0;
            if (++buf == 3) {
                buf = 0;
                ppar ^= 1u;
            }
        }
    } else {
        // ===== row warps: warp w owns TMEM lane quarter w % 4 (query rows) and the tiles of parity w / 4 =====
        const int q = warp & 3, hf = warp >> 2;
        const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
        auto read_out = [&](int j) {  // O of a finished head -> normalised head output (key half 0 only)
            const int ob = j & 1;
            mbar_wait(O_FULL(ob), (j >> 1) & 1);
            tc_fence_after();
            uint32_t o[8];
            tmem_ld8(trow + 192 + 16 * ob, o);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(O_READ(ob));
            const int qrow = t * 128 + 32 * q + lane;
            if (qrow >= L) return;
            const float inv = 1.0f / __uint_as_float(o[6]);
            if (att_img != nullptr) {
                // fp16 operand image of the out-proj / FFN kernel: per 256-token tile [kc][256 rows][8 halfs]
                const size_t mrow = (size_t)b * L + qrow;
                uint8_t *tile = reinterpret_cast<uint8_t *>(att_img) + (mrow >> 8) * (size_t)(9 * 256 * 16) + (mrow & 255) * 16;
                const int c0 = (g * HPC + j) * DH;
#pragma unroll
                for (int e = 0; e < 3; ++e) {
                    const int c = c0 + 2 * e;
                    *reinterpret_cast<uint32_t *>(tile + (c >> 3) * (256 * 16) + (c & 7) * 2) =
                        pack_f16x2_sat(__uint_as_float(o[2 * e + 1]) * inv, __uint_as_float(o[2 * e]) * inv);
                }
            } else {
                float *dst = att_out + ((size_t)b * L + qrow) * D + (g * HPC + j) * DH;
                reinterpret_cast<float2 *>(dst)[0] = make_float2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                reinterpret_cast<float2 *>(dst)[1] = make_float2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                reinterpret_cast<float2 *>(dst)[2] = make_float2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
            }
        };
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY, shift = 0.f;
        int pending = -1;  // head whose O still has to be read out (deferred into the next head so the read never waits)
        // Units alternate between the two warps that share my TMEM lanes: key half... rather UNIT parity hf — warp hf owns units u with
        // u % 2 == hf and processes the whole 64-key tile (two 32-key chunks) of its 32 query rows, so each warp synchronises with the
        // MMA warp once per TWO tiles and the two warps of a row work on different score buffers at the same time.
        int j = 0, kt = 0, buf = 0;  // the unit sequence as a cursor (no search per unit)
        bool mp = !bnd[0], bj = bnd[0];
        uint32_t spar = 0;
        for (int u = 0; u < U; ++u) {
            const bool mine = (u & 1) == hf;
            if (mine) {
                mbar_wait(S_FULL(buf), spar);
                tc_fence_after();
            }
            if (mp) {
                if (kt == 0) m0 = m1 = m2 = m3 = -INFINITY;
                if (mine) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int col = 64 * buf + 32 * c, key0 = kt * KT + 32 * c;
                        if (key0 + 32 <= L) max_chunk<false>(trow, col, L, m0, m1, m2, m3);
                        else if (key0 < L) max_chunk<true>(trow, col, L, m0, m1, m2, m3, key0);
                    }
                    tc_fence_before();
                    mbar_arrive_warp(P_READY(buf), lane);
                }
                if (kt == nkt - 1) {  // end of pass 1: combine with the thread that owns the other tiles of my row
                    float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    mx[hf * 128 + 32 * q + lane] = m;
                    pair_barrier_sync(q);
                    m = fmaxf(m, mx[(hf ^ 1) * 128 + 32 * q + lane]);
                    pair_barrier_sync(q);  // both have read before the next head overwrites the slots
                    shift = rintf(fminf(fmaxf(m, -4.0e6f), 4.0e6f));
                }
            } else {
                if (mine) {
                    // the whole 64-key tile as four 16-column sub-chunks, the next TMEM load in flight while one is exponentiated
                    const int col = 64 * buf, key0 = kt * KT;
                    if (bj) {
                        if (key0 + 64 <= L) exp_cols16_pipelined<false, 7, 16, 3, 4, 8, 32>(trow, col, col, key0, L, 0.f);
                        else exp_cols16_pipelined<true, 7, 16, 3, 4, 8, 32>(trow, col, col, key0, L, 0.f);
                    } else {
                        if (key0 + 64 <= L) exp_cols16_pipelined<false, 3, 8, 1, 4, 8, 32>(trow, col, col, key0, L, shift);
                        else exp_cols16_pipelined<true, 3, 8, 1, 4, 8, 32>(trow, col, col, key0, L, shift);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive_warp(P_READY(buf), lane);
                }
                if (hf == 0 && pending >= 0 && kt == (nkt > 3 ? 3 : nkt - 1)) {
                    read_out(pending);
                    pending = -1;
                }
                if (kt == nkt - 1) {
                    if (hf == 0 && pending >= 0) read_out(pending);  // (only when a head has very few key tiles)
                    pending = j;
                }
            }
            if (++kt == nkt) {
                kt = 0;
                if (mp) {
                    mp = false;
                } else {
                    ++j;
                    bj = j < HPC && bnd[j];
                    mp = j < HPC && !bj;
                }
            }
            if (++buf == 3) {
                buf = 0;
                spar ^= 1u;
            }
        }
        if (hf == 0 && pending >= 0) read_out(pending);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == ROW_WARPS + 1) tmem_dealloc(tmem, TMEM_COLS);
}

// ---- host side ------------------------------------------------------------------------------------------------------------------------
int attn_stream_supported(const fd_config &c) {
    return c.model_kind == FD_MODEL_TRANSFORMER && c.d_model == ats::D && c.n_head == ats::H && c.max_len > 256;
}

int attn_stream_finalize(fd_handle *h) {
    using namespace ats;
    for (auto &w : h->tl) {
        float *a = nullptr, *ab = nullptr;
        FD_CUDA(cudaMalloc((void **)&a, (size_t)2 * WP_BYTES));
        FD_CUDA(cudaMalloc((void **)&ab, (size_t)2 * NPW * sizeof(float)));
        h->owned.push_back(a);
        h->owned.push_back(ab);
        pack_qkv_half_weights_kernel<<<64, 256>>>(w.in_w, w.in_b, a, ab);
        FD_CUDA(cudaGetLastError());
        w.in_pack_half = a;
        w.in_bias_pack_half = ab;
    }
    FD_CUDA(cudaDeviceSynchronize());
    FD_CUDA(cudaFuncSetAttribute(qkv_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_PROJ));
    FD_CUDA(cudaFuncSetAttribute(attention_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ATT));
    return 0;
}

// att <- concat_heads softmax(q k^T / sqrt(dh)) v with q|k|v = in_proj(h) for max_len > 256: projection to operand images, then the
// streaming attention kernel.  att_img (nullable): write the FFN-layer kernel's fp16 operand image instead of fp32 rows to att_out.
int launch_attention_stream(fd_handle *h, int layer, const float *hbuf, float *att_out, void *att_img, int B, cudaStream_t s) {
    using namespace ats;
    const TransformerLayerW &w = h->tl[layer];
    FD_CHECK(w.in_pack_half && h->ws_qimg && h->ws_kvimg && h->ws_nrm, "launch_attention_stream: images / workspace missing");
    const int L = h->cfg.max_len, NT = (L + 127) / 128;
    const float qscale = (float)(1.4426950408889634 / sqrt((double)DH));
    const int bounded = h->attn_bounded;  // 0: always the exact two-pass softmax (fd_set_option / FD_ATTN_BOUNDED)
    FD_CUDA(cudaMemsetAsync(h->ws_nrm, 0, stream_nrm_words(B) * sizeof(unsigned), s));
    qkv_image_kernel<<<dim3(NT, B, 2), 160, SMEM_PROJ, s>>>(hbuf, w.in_pack_half, w.in_bias_pack_half, h->ws_qimg, h->ws_kvimg, h->ws_nrm, L, qscale);
    cudaError_t e = cudaGetLastError();
    FD_CHECK(e == cudaSuccess, "qkv_image_kernel launch failed: %s", cudaGetErrorString(e));
    attention_stream_kernel<<<dim3(NT, B, NG), THREADS, SMEM_ATT, s>>>(h->ws_qimg, h->ws_kvimg, h->ws_nrm, att_out, (__half *)att_img, L, bounded);
    e = cudaGetLastError();
    FD_CHECK(e == cudaSuccess, "attention_stream_kernel launch failed: %s", cudaGetErrorString(e));
    h->launches += 2;
    g_global_launches += 2;
    return 0;
}

}  // namespace fd
