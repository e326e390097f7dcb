// Softmax building blocks shared by the attention kernels: row maximum and P = 2^(s - max) of 32-key chunks straight out of TMEM.
#pragma once
#include "fd_tc.cuh"

namespace fd {

using namespace tc;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 64-thread named barrier 1 + q (immediate ids, so the kernel reserves 5 hardware barriers, not all 16 — they limit CTAs per SM)
__device__ __forceinline__ void pair_barrier_sync(int q) {
    switch (q) {
        case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    }
}

// 128-thread flavour: the four warps that share a TMEM lane quarter
__device__ __forceinline__ void quad_barrier_sync(int q) {
    switch (q) {
        case 0: asm volatile("bar.sync 1, 128;" ::: "memory"); break;
        case 1: asm volatile("bar.sync 2, 128;" ::: "memory"); break;
        case 2: asm volatile("bar.sync 3, 128;" ::: "memory"); break;
        default: asm volatile("bar.sync 4, 128;" ::: "memory"); break;
    }
}

// Softmax of one 128-query tile straight out of TMEM.  S columns = keys (already in log2 units); P = 2^(s - max) is written back as packed
// fp16 pairs — the A operand of a kind::f16 P·V MMA: quarter g (64 keys) reads S columns [64g, 64g+64) and leaves its 32 packed columns in
// [64g, 64g+32); columns [32, 48) — consumed with quarter 0 and never written again — hold the O accumulator.  fp16 carries the same 11
// significant bits as tf32; the halved exponent range is irrelevant for p in (0, 1].
//
// Instruction budget (measured on B200, tools/ubench/pipes.cu: clocks per warp instruction per SM sub-partition): MUFU.EX2 8 (and
// ex2.approx.f16x2 is TWO MUFU operations, 16), FMNMX / FMNMX3 2, FADD2 / FFMA2 2 (two fp32 lanes each).  The exponentials are therefore
// split between the MUFU pipe and a polynomial on the FMA pipe evaluated on packed fp32 pairs:
//     2^x, x <= 0:  x = n + f (round-to-nearest split by the magic-number trick), degree-3 minimax polynomial for 2^f on [-0.5, 0.5]
//     (relative error 1.0e-4, a fifth of the fp16 rounding that follows), n added into the exponent field.
// POLY_NUM of every POLY_DEN key pairs take the polynomial; the row maximum uses the 3-input FMNMX3 and the shift x = s - max one FADD2
// per pair.
constexpr int POLY_NUM = 1, POLY_DEN = 2;

__device__ __forceinline__ uint32_t exp2_pair_poly(uint64_t x2) {
    float x0, x1;
    f2_unpack(x2, x0, x1);
    x0 = fmaxf(x0, -25.0f);                     // 2^-25 rounds to fp16 zero; keeps n inside the fp32 exponent range
    x1 = fmaxf(x1, -25.0f);
    const uint64_t x = f2_pack(x0, x1);
    const uint64_t magic = f2_pack(12582912.0f, 12582912.0f);  // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const uint64_t t = f2_add(x, magic);
    const uint64_t f = f2_sub(x, f2_sub(t, magic));
    uint64_t p = f2_fma(f2_pack(0.05500893f, 0.05500893f), f, f2_pack(0.24221095f, 0.24221095f));
    p = f2_fma(p, f, f2_pack(0.6932829f, 0.6932829f));
    p = f2_fma(p, f, f2_pack(1.0f, 1.0f));
    float p0, p1, t0, t1;
    f2_unpack(p, p0, p1);
    f2_unpack(t, t0, t1);
    p0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
    return pack_f16x2(p1, p0);  // low half = even key
}

// pass 1, one 32-key chunk starting at S column `col`: fold into the running maxima; MASKED: keys >= L are ignored
// (the `_regs` forms work on a chunk that is already in registers, so a caller can prefetch the next chunk's tcgen05.ld)
template <bool MASKED>
__device__ __forceinline__ void max_regs(const uint32_t (&v)[32], int kb, int L, float &m0, float &m1, float &m2, float &m3) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        if (!MASKED || kb + j + 7 < L) {
            m0 = max3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            m1 = max3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            m2 = max3(m2, __uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
            m3 = max3(m3, __uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (kb + j + e < L) m0 = fmaxf(m0, __uint_as_float(v[j + e]));
        }
    }
}
template <bool MASKED>
__device__ __forceinline__ void max_chunk(uint32_t tS, int col, int L, float &m0, float &m1, float &m2, float &m3, int key0 = -1) {
    uint32_t v[32];
    tmem_ld32(tS + col, v);
    tmem_ld_wait();
    max_regs<MASKED>(v, key0 >= 0 ? key0 : col, L, m0, m1, m2, m3);
}

// 2^(s - c) for a pair of scores with the integer shift c folded into the magic constant: K = 1.5 * 2^23 - c is exact, t = s + K rounds to
// K + round(s) - c... so the low mantissa bits of t hold n = round(s) - c, t - K = round(s) exactly, and f = s - round(s) in [-0.5, 0.5].
// CLAMP: scores below c - 25 are raised to it first (their p rounds to fp16 zero either way; keeps n inside the fp32 exponent range).
template <bool CLAMP>
__device__ __forceinline__ uint32_t exp2_pair_poly_folded(float s0, float s1, uint64_t K2, float floor_s) {
    if (CLAMP) {
        s0 = fmaxf(s0, floor_s);
        s1 = fmaxf(s1, floor_s);
    }
    const uint64_t s2 = f2_pack(s0, s1);
    const uint64_t t = f2_add(s2, K2);
    const uint64_t f = f2_sub(s2, f2_sub(t, K2));
    uint64_t p = f2_fma(f2_pack(0.05500893f, 0.05500893f), f, f2_pack(0.24221095f, 0.24221095f));
    p = f2_fma(p, f, f2_pack(0.6932829f, 0.6932829f));
    p = f2_fma(p, f, f2_pack(1.0f, 1.0f));
    float p0, p1, t0, t1;
    f2_unpack(p, p0, p1);
    f2_unpack(t, t0, t1);
    p0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
    return pack_f16x2(p1, p0);  // low half = even key
}

// pass 2, one 32-key chunk: P = 2^(s - c) for S columns [col, col + 32) as 16 packed fp16 pairs into columns [pcol, pcol + 16).
// c = the row maximum rounded to an integer (softmax is invariant to the shift, every thread of the row uses the same c, p <= 2^0.5).
// VAR 0: x = s - c first, generic polynomial; 1: shift folded into the polynomial's magic constant; 2: as 1 without the underflow clamp
// (only legal when every score of the row is known to be > c - 120); 3: BOUNDED scores — the caller guarantees |s| <= 14 for the whole
// head, so P = 2^s needs no shift at all (softmax is shift-invariant, 2^-14 .. 2^14 are normal fp16 numbers) and no row maximum.
// key0: index of the chunk's first key (for masking) when it differs from the S column (rotating score buffers); < 0: the column is the key.
template <bool MASKED, int PN = POLY_NUM, int PD = POLY_DEN, int VAR = 1>
__device__ __forceinline__ void exp_regs(const uint32_t (&v)[32], uint32_t (&u)[16], int kb, int L, float c) {
    const uint64_t cc = f2_pack(c, c);
    const float Kf = 12582912.0f - c;
    const uint64_t K2 = f2_pack(Kf, Kf);
    const float floor_s = c - 25.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float e0 = __uint_as_float(v[2 * i]), e1 = __uint_as_float(v[2 * i + 1]);
        if (MASKED) {
            if (kb + 2 * i >= L) e0 = -INFINITY;
            if (kb + 2 * i + 1 >= L) e1 = -INFINITY;
        }
        if ((i * PN) % PD < PN) {
            if (VAR == 0) u[i] = exp2_pair_poly(f2_sub(f2_pack(e0, e1), cc));
            else u[i] = exp2_pair_poly_folded<VAR == 1 || MASKED>(e0, e1, K2, floor_s);
        } else if (VAR == 3) {
            u[i] = pack_f16x2(ex2_approx(e1), ex2_approx(e0));
        } else {
            float x0, x1;
            f2_unpack(f2_sub(f2_pack(e0, e1), cc), x0, x1);
            u[i] = pack_f16x2(ex2_approx(x1), ex2_approx(x0));  // low half = even key
        }
    }
}
template <bool MASKED, int PN = POLY_NUM, int PD = POLY_DEN, int VAR = 1>
__device__ __forceinline__ void exp_chunk(uint32_t tS, int col, int pcol, int L, float c, int key0 = -1) {
    uint32_t v[32], u[16];
    tmem_ld32(tS + col, v);
    tmem_ld_wait();
    exp_regs<MASKED, PN, PD, VAR>(v, u, key0 >= 0 ? key0 : col, L, c);
    tmem_st16(tS + pcol, u);
}

// Row maximum over NSUB 16-column sub-chunks starting at S column `col` (first key index kb), the next sub-chunk's tcgen05.ld in flight
// while one is folded in; keys >= L are ignored
template <int NSUB>
__device__ __forceinline__ void max_cols16_pipelined(uint32_t tS, int col, int kb, int L, float &m0, float &m1, float &m2, float &m3) {
    uint32_t va[16], vb[16];
    tmem_ld16(tS + col, va);
#pragma unroll
    for (int s = 0; s < NSUB; ++s) {
        uint32_t(&cur)[16] = (s & 1) ? vb : va;
        uint32_t(&nxt)[16] = (s & 1) ? va : vb;
        tmem_ld_wait16(cur);
        if (s + 1 < NSUB) tmem_ld16(tS + col + 16 * (s + 1), nxt);
        const int k0 = kb + 16 * s;
        if (k0 + 16 <= L) {
            m0 = max3(m0, __uint_as_float(cur[0]), __uint_as_float(cur[1]));
            m1 = max3(m1, __uint_as_float(cur[2]), __uint_as_float(cur[3]));
            m2 = max3(m2, __uint_as_float(cur[4]), __uint_as_float(cur[5]));
            m3 = max3(m3, __uint_as_float(cur[6]), __uint_as_float(cur[7]));
            m0 = max3(m0, __uint_as_float(cur[8]), __uint_as_float(cur[9]));
            m1 = max3(m1, __uint_as_float(cur[10]), __uint_as_float(cur[11]));
            m2 = max3(m2, __uint_as_float(cur[12]), __uint_as_float(cur[13]));
            m3 = max3(m3, __uint_as_float(cur[14]), __uint_as_float(cur[15]));
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
                if (k0 + e < L) m0 = fmaxf(m0, __uint_as_float(cur[e]));
        }
    }
}

// The same exponentials on 16-column sub-chunks with the NEXT sub-chunk's tcgen05.ld in flight while the current one is exponentiated
// (a row warp otherwise sits through the full TMEM load latency once per chunk; with four row warps per SM sub-partition that latency is
// not hidden by the other warps).  NSUB sub-chunks starting at S column `col`; sub-chunk s leaves its 8 packed P columns at
// pcol0 + PSTEP_LO * (s & 1) + PSTEP_HI * (s >> 1)  (fused kernel: 8, 16 — contiguous; streaming kernel: 8, 32 — [0,16) and [32,48)).
// The polynomial / MUFU pattern runs over the pairs of a whole 32-key chunk exactly as in exp_regs.
template <bool MASKED, int PN, int PD, int VAR, int NSUB, int PSTEP_LO, int PSTEP_HI>
__device__ __forceinline__ void exp_cols16_pipelined(uint32_t tS, int col, int pcol0, int kb, int L, float c) {
    const uint64_t cc = f2_pack(c, c);
    const float Kf = 12582912.0f - c;
    const uint64_t K2 = f2_pack(Kf, Kf);
    const float floor_s = c - 25.0f;
    uint32_t va[16], vb[16];
    tmem_ld16(tS + col, va);
#pragma unroll
    for (int s = 0; s < NSUB; ++s) {
        uint32_t(&cur)[16] = (s & 1) ? vb : va;
        uint32_t(&nxt)[16] = (s & 1) ? va : vb;
        tmem_ld_wait16(cur);
        if (s + 1 < NSUB) tmem_ld16(tS + col + 16 * (s + 1), nxt);
        uint32_t u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float e0 = __uint_as_float(cur[2 * i]), e1 = __uint_as_float(cur[2 * i + 1]);
            if (MASKED) {
                if (kb + 16 * s + 2 * i >= L) e0 = -INFINITY;
                if (kb + 16 * s + 2 * i + 1 >= L) e1 = -INFINITY;
            }
            const int ip = 8 * (s & 1) + i;  // pair index inside the 32-key chunk
            if ((ip * PN) % PD < PN) {
                if (VAR == 0) u[i] = exp2_pair_poly(f2_sub(f2_pack(e0, e1), cc));
                else u[i] = exp2_pair_poly_folded<VAR == 1 || MASKED>(e0, e1, K2, floor_s);
            } else if (VAR == 3) {
                u[i] = pack_f16x2(ex2_approx(e1), ex2_approx(e0));
            } else {
                float x0, x1;
                f2_unpack(f2_sub(f2_pack(e0, e1), cc), x0, x1);
                u[i] = pack_f16x2(ex2_approx(x1), ex2_approx(x0));
            }
        }
        tmem_st8(tS + pcol0 + PSTEP_LO * (s & 1) + PSTEP_HI * (s >> 1), u);
    }
}

}  // namespace fd
