// Register-resident DFT codelets for the dft / idft kernels (fd_fft.cu): Dft<R>::run(v) replaces v[0..R) by its forward DFT
// X[t] = sum_b v[b] exp(-2 pi i b t / R), natural order in and out.  Power-of-two sizes are hard-wired split-radix style butterflies on
// packed fp32 pairs (one FADD2 per complex add); odd sizes use the conjugate-pair form
//     X[t], X[R-t] = v0 + sum_b (v_b + v_{R-b}) cos(2 pi b t / R)  -/+  i sum_b (v_b - v_{R-b}) sin(2 pi b t / R),   b = 1 .. (R-1)/2
// (half the multiplications of the plain O(R^2) sum) with the cos / sin values in constant memory, where an FFMA reads them as operands.
#pragma once
#include <cuda_runtime.h>

namespace fd {
namespace fft {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// complex add / subtract as ONE packed fp32x2 instruction (FADD2): the butterflies are issue-bound, not flop-bound
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // * -i
// a * (c - i s): the forward twiddle exp(-i theta) with c = cos(theta), s = sin(theta)
__device__ __forceinline__ float2 mul_cs(float2 a, float c, float s) { return make_float2(fmaf(a.y, s, a.x * c), fmaf(-a.x, s, a.y * c)); }

__constant__ float2 k_trig_3[3] = {{1.000000000e+00f, 0.000000000e+00f}, {-5.000000000e-01f, 8.660254038e-01f}, {-5.000000000e-01f, -8.660254038e-01f}};
__constant__ float2 k_trig_5[5] = {{1.000000000e+00f, 0.000000000e+00f}, {3.090169944e-01f, 9.510565163e-01f}, {-8.090169944e-01f, 5.877852523e-01f}, {-8.090169944e-01f, -5.877852523e-01f}, {3.090169944e-01f, -9.510565163e-01f}};
__constant__ float2 k_trig_7[7] = {{1.000000000e+00f, 0.000000000e+00f}, {6.234898019e-01f, 7.818314825e-01f}, {-2.225209340e-01f, 9.749279122e-01f}, {-9.009688679e-01f, 4.338837391e-01f}, {-9.009688679e-01f, -4.338837391e-01f}, {-2.225209340e-01f, -9.749279122e-01f}, {6.234898019e-01f, -7.818314825e-01f}};
__constant__ float2 k_trig_9[9] = {{1.000000000e+00f, 0.000000000e+00f}, {7.660444431e-01f, 6.427876097e-01f}, {1.736481777e-01f, 9.848077530e-01f}, {-5.000000000e-01f, 8.660254038e-01f}, {-9.396926208e-01f, 3.420201433e-01f}, {-9.396926208e-01f, -3.420201433e-01f}, {-5.000000000e-01f, -8.660254038e-01f}, {1.736481777e-01f, -9.848077530e-01f}, {7.660444431e-01f, -6.427876097e-01f}};
__constant__ float2 k_trig_11[11] = {{1.000000000e+00f, 0.000000000e+00f}, {8.412535328e-01f, 5.406408175e-01f}, {4.154150130e-01f, 9.096319954e-01f}, {-1.423148383e-01f, 9.898214419e-01f}, {-6.548607339e-01f, 7.557495744e-01f}, {-9.594929736e-01f, 2.817325568e-01f}, {-9.594929736e-01f, -2.817325568e-01f}, {-6.548607339e-01f, -7.557495744e-01f}, {-1.423148383e-01f, -9.898214419e-01f}, {4.154150130e-01f, -9.096319954e-01f}, {8.412535328e-01f, -5.406408175e-01f}};
__constant__ float2 k_trig_13[13] = {{1.000000000e+00f, 0.000000000e+00f}, {8.854560257e-01f, 4.647231720e-01f}, {5.680647467e-01f, 8.229838659e-01f}, {1.205366803e-01f, 9.927088741e-01f}, {-3.546048870e-01f, 9.350162427e-01f}, {-7.485107482e-01f, 6.631226582e-01f}, {-9.709418174e-01f, 2.393156643e-01f}, {-9.709418174e-01f, -2.393156643e-01f}, {-7.485107482e-01f, -6.631226582e-01f}, {-3.546048870e-01f, -9.350162427e-01f}, {1.205366803e-01f, -9.927088741e-01f}, {5.680647467e-01f, -8.229838659e-01f}, {8.854560257e-01f, -4.647231720e-01f}};
__constant__ float2 k_trig_17[17] = {{1.000000000e+00f, 0.000000000e+00f}, {9.324722294e-01f, 3.612416662e-01f}, {7.390089172e-01f, 6.736956436e-01f}, {4.457383558e-01f, 8.951632914e-01f}, {9.226835946e-02f, 9.957341763e-01f}, {-2.736629901e-01f, 9.618256432e-01f}, {-6.026346364e-01f, 7.980172273e-01f}, {-8.502171357e-01f, 5.264321629e-01f}, {-9.829730997e-01f, 1.837495178e-01f}, {-9.829730997e-01f, -1.837495178e-01f}, {-8.502171357e-01f, -5.264321629e-01f}, {-6.026346364e-01f, -7.980172273e-01f}, {-2.736629901e-01f, -9.618256432e-01f}, {9.226835946e-02f, -9.957341763e-01f}, {4.457383558e-01f, -8.951632914e-01f}, {7.390089172e-01f, -6.736956436e-01f}, {9.324722294e-01f, -3.612416662e-01f}};

template <int R>
__device__ __forceinline__ float2 trig(int m) {  // (cos, sin)(2 pi m / R); m is a compile-time constant after unrolling
    static_assert(R == 3 || R == 5 || R == 7 || R == 9 || R == 11 || R == 13 || R == 17, "odd radix without a table");
    return R == 3 ? k_trig_3[m] : R == 5 ? k_trig_5[m] : R == 7 ? k_trig_7[m] : R == 9 ? k_trig_9[m] : R == 11 ? k_trig_11[m] : R == 13 ? k_trig_13[m] : k_trig_17[m];
}

template <int R>
struct Dft {  // odd R
    static __device__ __forceinline__ void run(float2 (&v)[R]) {
        constexpr int H = (R - 1) / 2;
        float2 s[H], d[H];
#pragma unroll
        for (int b = 1; b <= H; ++b) {
            s[b - 1] = cadd(v[b], v[R - b]);
            d[b - 1] = csub(v[b], v[R - b]);
        }
        const float2 v0 = v[0];
        float2 x0 = v0;
#pragma unroll
        for (int b = 0; b < H; ++b) x0 = cadd(x0, s[b]);
        v[0] = x0;
#pragma unroll
        for (int t = 1; t <= H; ++t) {
            float2 A = v0, Bs = make_float2(0.f, 0.f);
#pragma unroll
            for (int b = 1; b <= H; ++b) {
                const float2 cs = trig<R>((b * t) % R);
                A.x = fmaf(s[b - 1].x, cs.x, A.x);
                A.y = fmaf(s[b - 1].y, cs.x, A.y);
                Bs.x = fmaf(d[b - 1].x, cs.y, Bs.x);
                Bs.y = fmaf(d[b - 1].y, cs.y, Bs.y);
            }
            // X[t] = A - i Bs,  X[R - t] = A + i Bs
            v[t] = make_float2(A.x + Bs.y, A.y - Bs.x);
            v[R - t] = make_float2(A.x - Bs.y, A.y + Bs.x);
        }
    }
};
template <>
struct Dft<1> {
    static __device__ __forceinline__ void run(float2 (&)[1]) {}
};
template <>
struct Dft<2> {
    static __device__ __forceinline__ void run(float2 (&v)[2]) {
        const float2 a = cadd(v[0], v[1]), b = csub(v[0], v[1]);
        v[0] = a;
        v[1] = b;
    }
};
template <>
struct Dft<4> {
    static __device__ __forceinline__ void run(float2 (&v)[4]) {
        const float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
        const float2 a2 = cadd(v[1], v[3]), a3 = mul_mi(csub(v[1], v[3]));
        v[0] = cadd(a0, a2);
        v[1] = cadd(a1, a3);
        v[2] = csub(a0, a2);
        v[3] = csub(a1, a3);
    }
};
template <>
struct Dft<8> {
    static __device__ __forceinline__ void run(float2 (&v)[8]) {
        const float h = 0.70710678118654752f;
        float2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
        Dft<4>::run(e);
        Dft<4>::run(o);
        o[1] = make_float2((o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h);   // * (1 - i) / sqrt 2
        o[2] = mul_mi(o[2]);
        o[3] = make_float2((o[3].y - o[3].x) * h, -(o[3].x + o[3].y) * h);  // * (-1 - i) / sqrt 2
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = cadd(e[k], o[k]);
            v[k + 4] = csub(e[k], o[k]);
        }
    }
};
template <>
struct Dft<16> {
    static __device__ __forceinline__ void run(float2 (&v)[16]) {
        const float h = 0.70710678118654752f, c1 = 0.92387953251128674f, s1 = 0.38268343236508977f;
        float2 e[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            e[k] = v[2 * k];
            o[k] = v[2 * k + 1];
        }
        Dft<8>::run(e);
        Dft<8>::run(o);
        o[1] = mul_cs(o[1], c1, s1);
        o[2] = make_float2((o[2].x + o[2].y) * h, (o[2].y - o[2].x) * h);
        o[3] = mul_cs(o[3], s1, c1);
        o[4] = mul_mi(o[4]);
        o[5] = mul_cs(o[5], -s1, c1);
        o[6] = make_float2((o[6].y - o[6].x) * h, -(o[6].x + o[6].y) * h);
        o[7] = mul_cs(o[7], -c1, s1);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v[k] = cadd(e[k], o[k]);
            v[k + 8] = csub(e[k], o[k]);
        }
    }
};

}  // namespace fft
}  // namespace fd
