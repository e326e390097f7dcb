// dft / idft of the fdiff sampler (src/fdiff/utils/fourier.py:8-87): ortho real FFT along dim 1 of a (B, L, C) tensor with
// the reference's packed-real layout  [Re X_0..X_{L/2} | Im X_1..X_{ceil(L/2)-1}].
//
// One CTA owns one series (all channels, or a group of channel pairs when the slab does not fit in shared memory).  The
// (L, C) slab is read once with coalesced loads, two real channels are packed into one complex sequence (z = x_a + i x_b),
// a mixed-radix Stockham FFT (any L: radix 4/2/3/5/7 and arbitrary prime radices) runs entirely in shared memory with an
// exact twiddle table exp(-2*pi*i*q/L) computed in fp64 on the host, and the result is unpacked / written once.
// Algorithmic HBM traffic: 8*B*L*C bytes per transform (read + write once).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include <map>
#include <mutex>
#include <vector>

#include "fd_common.cuh"
#include "fd_fft_codelets.cuh"
#include "fd_tc.cuh"

namespace fd {

using fft::cadd;
using fft::cmul;
using fft::csub;

constexpr int FFT_MAX_STAGES = 24;
struct FftPlan {
    int n_stages;
    int radix[FFT_MAX_STAGES];
};

// One Stockham stage, thread per output element.  src/dst: [L][P] complex.  tw: exp(-2 pi i q / L) (conjugated on the fly
// for the inverse).  Ns = product of the radices of earlier stages.
__device__ __forceinline__ void stockham_stage(const float2 *__restrict__ src, float2 *__restrict__ dst, const float2 *__restrict__ tw,
                                               int L, int P, int R, int Ns, bool inverse) {
    const int LR = L / R;
    const int tstride = L / (Ns * R);
    for (int idx = threadIdx.x; idx < L * P; idx += blockDim.x) {
        int p = idx % P, o = idx / P;
        int k = o % Ns;
        int t = (o / Ns) % R;
        int jhi = o / (Ns * R);
        int j = jhi * Ns + k;
        int qstep = k * tstride + t * LR;
        if (qstep >= L) qstep -= L;
        int q = 0;
        float2 acc = make_float2(0.f, 0.f);
        const float2 *in = src + (size_t)j * P + p;
        for (int b = 0; b < R; ++b) {
            float2 w = tw[q];
            if (inverse) w.y = -w.y;
            float2 v = in[(size_t)b * LR * P];
            acc.x = fmaf(v.x, w.x, acc.x);
            acc.x = fmaf(-v.y, w.y, acc.x);
            acc.y = fmaf(v.x, w.y, acc.y);
            acc.y = fmaf(v.y, w.x, acc.y);
            q += qstep;
            if (q >= L) q -= L;
        }
        dst[idx] = acc;
    }
}

// One Stockham stage, thread per radix-R BUTTERFLY (R inputs -> R outputs; used for R <= 8; forward transform only — the inverse runs it on re/im-swapped data): the inputs are multiplied by their stage
// twiddles W_L^(b k tstride) once, then an R-point DFT (hard-wired for R = 2 and 4, table-driven otherwise) produces all R outputs, so
// every input is loaded once per stage instead of R times.  src/dst: [L][NP] complex (NP sequences side by side).
// j -> (j / Ns, j % Ns) without an integer division: shift / mask for power-of-two Ns, else a float reciprocal with a one-step fix-up
// (exact for j < 2^22; j < 8192 here)
__device__ __forceinline__ void split_index(int j, int Ns, int log2Ns, float invNs, int &hi, int &lo) {
    if (log2Ns >= 0) {
        hi = j >> log2Ns;
        lo = j & (Ns - 1);
    } else {
        hi = (int)((float)j * invNs);
        lo = j - hi * Ns;
        if (lo >= Ns) {
            lo -= Ns;
            ++hi;
        } else if (lo < 0) {
            lo += Ns;
            --hi;
        }
    }
}

// e -> e / P for 0 <= e < 2^22 without an integer division (float reciprocal, one-step fix-up)
__device__ __forceinline__ int fast_div(int e, int P, float invP) {
    int q = (int)((float)e * invP);
    const int r = e - q * P;
    if (r >= P) ++q;
    else if (r < 0) --q;
    return q;
}

template <int R>
__device__ __forceinline__ void butterfly_stage(const float2 *__restrict__ src, float2 *__restrict__ dst, const float2 *__restrict__ tw, int L,
                                                int NP, int Ns, bool inverse) {
    const int LR = L / R;
    const int tstride = L / (Ns * R);
    const int log2Ns = (Ns & (Ns - 1)) == 0 ? 31 - __clz(Ns) : -1;
    const float invNs = 1.0f / (float)Ns;
    const int os = Ns * NP, in_step = LR * NP;
    // (j, p) of my first butterfly and the per-iteration step: no division inside the loop
    int j = threadIdx.x / NP, p = threadIdx.x - j * NP;
    const int dj = blockDim.x / NP, dp = blockDim.x - dj * NP;
    for (; j < LR;) {
        int jhi, k;
        split_index(j, Ns, log2Ns, invNs, jhi, k);
        float2 v[R];
        const float2 *in = src + j * NP + p;
#pragma unroll
        for (int b = 0; b < R; ++b) v[b] = in[b * in_step];
        const int q1 = k * tstride;  // < L / R, so b * q1 < L
#pragma unroll
        for (int b = 1; b < R; ++b) {
            float2 w = tw[b * q1];
            if (inverse) w.y = -w.y;
            v[b] = cmul(v[b], w);
        }
        float2 *o = dst + (jhi * R * Ns + k) * NP + p;
        if (R == 2) {
            o[0] = cadd(v[0], v[1]);
            o[os] = csub(v[0], v[1]);
        } else if (R == 4) {
            const float2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
            const float2 a2 = cadd(v[1], v[3]), a3 = csub(v[1], v[3]);
            // forward: W_4 = -i;  inverse: +i
            const float2 ja3 = inverse ? make_float2(-a3.y, a3.x) : make_float2(a3.y, -a3.x);
            o[0] = cadd(a0, a2);
            o[os] = cadd(a1, ja3);
            o[2 * os] = csub(a0, a2);
            o[3 * os] = csub(a1, ja3);
        } else if (R == 8) {
            // forward DFT-8 (the inverse transform runs the forward FFT on re/im-swapped data): split into even / odd DFT-4s
            const float h = 0.70710678118654752f;
            float2 a[4], c[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = cadd(v[i], v[i + 4]);
                c[i] = csub(v[i], v[i + 4]);
            }
            c[1] = make_float2((c[1].x + c[1].y) * h, (c[1].y - c[1].x) * h);    // * (1 - i)/sqrt2
            c[2] = make_float2(c[2].y, -c[2].x);                                // * -i
            c[3] = make_float2((c[3].y - c[3].x) * h, -(c[3].x + c[3].y) * h);  // * (-1 - i)/sqrt2
#pragma unroll
            for (int par = 0; par < 2; ++par) {
                const float2 *xin = par ? c : a;
                const float2 s0 = cadd(xin[0], xin[2]), s1 = csub(xin[0], xin[2]);
                const float2 s2 = cadd(xin[1], xin[3]), d3 = csub(xin[1], xin[3]);
                const float2 s3 = make_float2(d3.y, -d3.x);  // * -i
                o[(0 + par) * os] = cadd(s0, s2);
                o[(2 + par) * os] = cadd(s1, s3);
                o[(4 + par) * os] = csub(s0, s2);
                o[(6 + par) * os] = csub(s1, s3);
            }
        } else {
#pragma unroll
            for (int t = 0; t < R; ++t) {
                float2 acc = v[0];
#pragma unroll
                for (int b = 1; b < R; ++b) {
                    float2 w = tw[((b * t) % R) * LR];  // W_R^(b t)
                    if (inverse) w.y = -w.y;
                    acc.x = fmaf(v[b].x, w.x, acc.x);
                    acc.x = fmaf(-v[b].y, w.y, acc.x);
                    acc.y = fmaf(v[b].x, w.y, acc.y);
                    acc.y = fmaf(v[b].y, w.x, acc.y);
                }
                o[t * os] = acc;
            }
        }
        j += dj;
        p += dp;
        if (p >= NP) {
            p -= NP;
            ++j;
        }
    }
}

// Radix-8 Stockham stage IN PLACE: every thread first pulls the inputs of all its (up to NB) butterflies into registers, the CTA
// synchronises, then the butterflies are computed and written back into the SAME buffer.  Halves the shared memory of a transform, which is
// what lets four channel pairs of an L = 4096 series (instead of two) share a CTA — whole 32-byte sectors per row instead of half ones.
template <int NB>
__device__ __forceinline__ void radix8_stage_inplace(float2 *__restrict__ buf, const float2 *__restrict__ tw, int L, int NP, int Ns) {
    constexpr int R = 8;
    const int LR = L / R, tstride = L / (Ns * R), total = LR * NP;
    const int log2Ns = (Ns & (Ns - 1)) == 0 ? 31 - __clz(Ns) : -1;
    const float invNs = 1.0f / (float)Ns, invNP = 1.0f / (float)NP;
    const int os = Ns * NP, in_step = LR * NP;
    float2 v[NB][R];
    int obase[NB], q1[NB];
#pragma unroll
    for (int n = 0; n < NB; ++n) {
        const int idx = threadIdx.x + n * blockDim.x;
        obase[n] = -1;
        if (idx < total) {
            const int j = fast_div(idx, NP, invNP), p = idx - j * NP;
            int jhi, k;
            split_index(j, Ns, log2Ns, invNs, jhi, k);
            const float2 *in = buf + j * NP + p;
#pragma unroll
            for (int b = 0; b < R; ++b) v[n][b] = in[b * in_step];
            obase[n] = (jhi * R * Ns + k) * NP + p;
            q1[n] = k * tstride;
        }
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < NB; ++n) {
        if (obase[n] < 0) continue;
#pragma unroll
        for (int b = 1; b < R; ++b) v[n][b] = cmul(v[n][b], tw[b * q1[n]]);
        float2 *o = buf + obase[n];
        const float h = 0.70710678118654752f;
        float2 a[4], c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = cadd(v[n][i], v[n][i + 4]);
            c[i] = csub(v[n][i], v[n][i + 4]);
        }
        c[1] = make_float2((c[1].x + c[1].y) * h, (c[1].y - c[1].x) * h);    // * (1 - i)/sqrt2
        c[2] = make_float2(c[2].y, -c[2].x);                                // * -i
        c[3] = make_float2((c[3].y - c[3].x) * h, -(c[3].x + c[3].y) * h);  // * (-1 - i)/sqrt2
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            const float2 *xin = par ? c : a;
            const float2 s0 = cadd(xin[0], xin[2]), s1 = csub(xin[0], xin[2]);
            const float2 s2 = cadd(xin[1], xin[3]), d3 = csub(xin[1], xin[3]);
            const float2 s3 = make_float2(d3.y, -d3.x);  // * -i
            o[(0 + par) * os] = cadd(s0, s2);
            o[(2 + par) * os] = cadd(s1, s3);
            o[(4 + par) * os] = csub(s0, s2);
            o[(6 + par) * os] = csub(s1, s3);
        }
    }
}

// grid: (ceil(B / S), n_groups); each CTA handles S consecutive series and channel pairs [g*Pc, min(P, (g+1)*Pc)) of each: NP = S * P
// complex sequences side by side in shared memory ([L][NP]).  Global traffic is one coalesced pass in and one out (float2 accesses when
// the channel count is even).
template <bool INPLACE>
__global__ void __launch_bounds__(512) rfft_packed_kernel(const float *__restrict__ x, float *__restrict__ out,
                                                          const float2 *__restrict__ tw_g, FftPlan plan, int B, int L, int C, int Pc, int S,
                                                          const float *__restrict__ mean, const float *__restrict__ stdv, int inverse) {
    constexpr bool inplace = INPLACE;
    extern __shared__ float2 fsm[];
    float2 *tw = fsm;            // [L]
    const int b0 = blockIdx.x * S;
    const int nser = min(S, B - b0);
    const int p0 = blockIdx.y * Pc;
    const int Ptot = (C + 1) / 2;
    const int P = min(Pc, Ptot - p0);
    const int NP = nser * P;
    float2 *buf0 = tw + L;       // [L][NP]
    float2 *buf1 = inplace ? buf0 : buf0 + (size_t)L * S * Pc;  // in-place plans (all radix 8) run in one buffer
    const int n_real = L / 2 + 1;  // == ceil((L+1)/2), fourier.py:59
    const float scale = 1.0f / sqrtf((float)L);
    const bool vec2 = (C & 1) == 0;  // channel pairs are 8-byte aligned float2s
    const float invP = 1.0f / (float)P;

    // Forward transform of a whole, even-channel series group: the (L, C) slab in global memory IS the packed complex layout [L][P]
    // (z_p[l] = x[l][2p] + i x[l][2p+1]), so one bulk async copy (TMA engine) stages it — no load instructions, no registers in flight.
    const bool bulk = !inverse && vec2 && S == 1 && P == Ptot && (L & 1) == 0 && ((size_t)L * C * 4) % 16 == 0 && (reinterpret_cast<size_t>(x) & 15) == 0;
    const uint32_t bar = tc::smem_u32(buf1 + (size_t)L * S * Pc);  // 8 bytes behind the buffer(s)
    if (bulk && threadIdx.x == 0) {
        tc::mbar_init(bar, 1);
        tc::mbar_fence_init();
        tc::mbar_arrive_expect_tx(bar, (uint32_t)((size_t)L * C * 4));
        tc::bulk_g2s(tc::smem_u32(buf0), x + (size_t)b0 * L * C, (uint32_t)((size_t)L * C * 4), bar);
    }
    for (int i = threadIdx.x; i < L; i += blockDim.x) tw[i] = tw_g[i];

    constexpr int U = 4;  // independent global loads in flight per thread
    for (int si = 0; si < nser && !bulk; ++si) {
        const float *xs = x + (size_t)(b0 + si) * L * C;
        if (!inverse) {
            // load: z_p[l] = x[l][2p] + i x[l][2p+1]
            const int total = L * P;
            for (int e0 = threadIdx.x; e0 < total; e0 += U * blockDim.x) {
                float2 z[U];
                int l[U], pp[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = e0 + u * blockDim.x;
                    l[u] = fast_div(e, P, invP);
                    pp[u] = e - l[u] * P;
                    if (e < total) {
                        const int c0 = 2 * (p0 + pp[u]);
                        if (vec2) {
                            z[u] = *reinterpret_cast<const float2 *>(xs + (size_t)l[u] * C + c0);
                        } else {
                            z[u].x = xs[(size_t)l[u] * C + c0];
                            z[u].y = (c0 + 1 < C) ? xs[(size_t)l[u] * C + c0 + 1] : 0.f;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (e0 + u * blockDim.x < total) buf0[(size_t)l[u] * NP + si * P + pp[u]] = z[u];
            }
        } else {
            // rebuild the full spectrum of both channels from the packed layout (fourier.py:59-76), de-standardised first
            // (cmd/sample.py:76-78), and pack  Z[k] = X_a[k] + i X_b[k]; row kk of the input feeds both Z[kk] and Z[L - kk]
            const int total = n_real * P;
            for (int e0 = threadIdx.x; e0 < total; e0 += U * blockDim.x) {
                float2 xr[U], xi[U];
                int kk[U], pp[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = e0 + u * blockDim.x;
                    kk[u] = fast_div(e, P, invP);
                    pp[u] = e - kk[u] * P;
                    xr[u] = xi[u] = make_float2(0.f, 0.f);
                    if (e < total) {
                        const bool has_im = !(kk[u] == 0 || (L % 2 == 0 && kk[u] == L / 2));
                        const int c0 = 2 * (p0 + pp[u]);
                        const size_t ir = (size_t)kk[u] * C + c0, ii = (size_t)(n_real + kk[u] - 1) * C + c0;
                        if (vec2) {
                            xr[u] = *reinterpret_cast<const float2 *>(xs + ir);
                            if (has_im) xi[u] = *reinterpret_cast<const float2 *>(xs + ii);
                            if (mean) {
                                const float2 sr = *reinterpret_cast<const float2 *>(stdv + ir), mr = *reinterpret_cast<const float2 *>(mean + ir);
                                xr[u].x = xr[u].x * sr.x + mr.x;
                                xr[u].y = xr[u].y * sr.y + mr.y;
                                if (has_im) {
                                    const float2 s2 = *reinterpret_cast<const float2 *>(stdv + ii), m2 = *reinterpret_cast<const float2 *>(mean + ii);
                                    xi[u].x = xi[u].x * s2.x + m2.x;
                                    xi[u].y = xi[u].y * s2.y + m2.y;
                                }
                            }
                        } else {
                            xr[u].x = xs[ir];
                            if (mean) xr[u].x = xr[u].x * stdv[ir] + mean[ir];
                            if (has_im) {
                                xi[u].x = xs[ii];
                                if (mean) xi[u].x = xi[u].x * stdv[ii] + mean[ii];
                            }
                            if (c0 + 1 < C) {
                                xr[u].y = xs[ir + 1];
                                if (mean) xr[u].y = xr[u].y * stdv[ir + 1] + mean[ir + 1];
                                if (has_im) {
                                    xi[u].y = xs[ii + 1];
                                    if (mean) xi[u].y = xi[u].y * stdv[ii + 1] + mean[ii + 1];
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (e0 + u * blockDim.x < total) {
                        float2 *col = buf0 + si * P + pp[u];
                        // stored with re / im SWAPPED: the inverse transform is the forward FFT of the swapped data, swapped back on output
                        col[(size_t)kk[u] * NP] = make_float2(xi[u].x + xr[u].y, xr[u].x - xi[u].y);
                        if (kk[u] > 0 && 2 * kk[u] != L)  // mirror bin L - kk: conjugate spectra
                            col[(size_t)(L - kk[u]) * NP] = make_float2(xr[u].y - xi[u].x, xr[u].x + xi[u].y);
                    }
                }
            }
        }
    }
    __syncthreads();  // (also orders the mbarrier initialisation before everybody's wait)
    if (bulk) tc::mbar_wait(bar, 0);

    float2 *src = buf0, *dst = buf1;
    int Ns = 1;
    for (int s = 0; s < plan.n_stages; ++s) {
        const int R = plan.radix[s];
        if (inplace) {
            radix8_stage_inplace<4>(buf0, tw, L, NP, Ns);
            Ns *= R;
            __syncthreads();
            continue;
        }
        switch (R) {
            case 2: butterfly_stage<2>(src, dst, tw, L, NP, Ns, false); break;
            case 3: butterfly_stage<3>(src, dst, tw, L, NP, Ns, false); break;
            case 4: butterfly_stage<4>(src, dst, tw, L, NP, Ns, false); break;
            case 5: butterfly_stage<5>(src, dst, tw, L, NP, Ns, false); break;
            case 7: butterfly_stage<7>(src, dst, tw, L, NP, Ns, false); break;
            case 8: butterfly_stage<8>(src, dst, tw, L, NP, Ns, false); break;
            default: stockham_stage(src, dst, tw, L, NP, R, Ns, false); break;  // large prime factor: thread per output
        }
        Ns *= R;
        __syncthreads();
        float2 *t = src;
        src = dst;
        dst = t;
    }

    for (int si = 0; si < nser; ++si) {
        float *os = out + (size_t)(b0 + si) * L * C;
        if (!inverse) {
            // unpack the two real spectra and write the packed-real layout (fourier.py:21-40)
            for (int e = threadIdx.x; e < n_real * P; e += blockDim.x) {
                const int k = fast_div(e, P, invP), pp = e - k * P;
                const float2 zk = src[(size_t)k * NP + si * P + pp];
                const float2 zn = src[(size_t)((L - k) % L) * NP + si * P + pp];
                // X_a = (Z[k] + conj(Z[L-k]))/2 ; X_b = (Z[k] - conj(Z[L-k]))/(2i)
                const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
                const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
                const bool has_im = !(k == 0 || (L % 2 == 0 && k == L / 2));
                const int c0 = 2 * (p0 + pp);
                if (vec2) {
                    *reinterpret_cast<float2 *>(os + (size_t)k * C + c0) = make_float2(ar * scale, br * scale);
                    if (has_im) *reinterpret_cast<float2 *>(os + (size_t)(n_real + k - 1) * C + c0) = make_float2(ai * scale, bi * scale);
                } else {
                    os[(size_t)k * C + c0] = ar * scale;
                    if (has_im) os[(size_t)(n_real + k - 1) * C + c0] = ai * scale;
                    if (c0 + 1 < C) {
                        os[(size_t)k * C + c0 + 1] = br * scale;
                        if (has_im) os[(size_t)(n_real + k - 1) * C + c0 + 1] = bi * scale;
                    }
                }
            }
        } else {
            int l = threadIdx.x / P, pp = threadIdx.x - l * P;
            const int dl = blockDim.x / P, dpp = blockDim.x - dl * P;
            for (; l < L; l += dl, pp += dpp) {
                if (pp >= P) {
                    pp -= P;
                    if (++l >= L) break;
                }
                const float2 zs = src[(size_t)l * NP + si * P + pp];
                const float2 z = make_float2(zs.y, zs.x);  // swap back (see the load)
                const int c0 = 2 * (p0 + pp);
                if (vec2) {
                    *reinterpret_cast<float2 *>(os + (size_t)l * C + c0) = make_float2(z.x * scale, z.y * scale);
                } else {
                    os[(size_t)l * C + c0] = z.x * scale;
                    if (c0 + 1 < C) os[(size_t)l * C + c0 + 1] = z.y * scale;
                }
            }
        }
    }
}


// =======================================================================================================================================
// Fast paths.  Both pack the SAME column of two consecutive SERIES into one complex sequence (z = x_{2s} + i x_{2s+1}) instead of two
// channels of one series: the (L, C) slab of a series is then a flat array whose element (l, c) sits at l * C + c for ANY channel count,
// threads run over that flat index, and every global access of a warp is one contiguous run of 4-byte words — odd C, C = 1 included.
//
// rfft_cols_kernel: Stockham stages with the butterflies in registers (radix 16 / 8 / 4 / 2 and odd radices up to 17, fd_fft_codelets.cuh),
// one butterfly per thread and stage, IN PLACE in shared memory (all inputs of a stage are in registers before anything is written).  The
// first stage of the forward transform reads global memory directly and the last stage of the inverse writes it directly, so a 256-point
// transform makes two passes (16 x 16) with one shared-memory exchange, a 4096-point one three.  Twiddles come from the L1-resident table.
// rfft_small_kernel: max_len <= 32 — a whole (series pair, column) sequence per thread, no shared memory at all.
// =======================================================================================================================================
struct ColPlan {
    int n_stages;
    int radix[4];
    int L, C, CW, SP, Jmax;  // CW columns of SP series pairs per CTA; Jmax = butterflies per column of the widest stage
    int Lseq;                // Bluestein only: the series length (L is then the power-of-two convolution length M >= 2 Lseq - 1)
};

// MODE 0: shared -> shared in place; 1: global -> shared (first stage of the forward transform); 2: shared -> global (last stage of the inverse);
// 3 / 4: as 1 / 2 for the channel-pair packing (z = x[:, 2c] + i x[:, 2c+1] of ONE series: a complex element is one aligned float2 in global memory)
// SWZ (rows of 32 bytes, CW = 4): row r lives in slot r ^ ((r >> 4) & 3) — the radix-16 first stage writes rows 16 j + t from consecutive j, i.e.
// 512 bytes apart: without the swizzle the 8 rows of a warp store fall on the same 8 banks (8-way conflict), with it on all 32 (the 2-wavefront minimum).
// TWP (R = 16): the 15 twiddles W^(b q) from TWO loaded ones (W^q, W^4q) by products of depth <= 3 — the kernel is bound by the L1 / shared
// data pipe (ncu: 67 % wavefront utilisation, two thirds of the global-load sectors were twiddles), the fp32 pipes idle.
template <bool SWZ>
__device__ __forceinline__ int col_slot(int r) { return SWZ ? (r ^ ((r >> 4) & 3)) : r; }

template <int R, int MODE, bool SWZ = false, bool TWP = false>
__device__ __forceinline__ void col_stage(float2 *__restrict__ sb, const int CW, const int L, const int Ns, const int j, const bool thread_on,
                                          const float2 *__restrict__ tw, const float *__restrict__ ga, const float *__restrict__ gb,
                                          float *__restrict__ oa, float *__restrict__ ob, const int C, const float scale) {
    const int LR = L / R;
    const bool act = thread_on && j < LR;
    float2 v[R];
    if (act) {
        if (MODE == 1) {
#pragma unroll
            for (int b = 0; b < R; ++b) {
                const int row = j + b * LR;
                v[b].x = ga[row * C];
                v[b].y = gb ? gb[row * C] : 0.f;
            }
        } else if (MODE == 3) {
#pragma unroll
            for (int b = 0; b < R; ++b) v[b] = __ldcs(reinterpret_cast<const float2 *>(ga + (j + b * LR) * C));
        } else {
#pragma unroll
            for (int b = 0; b < R; ++b) v[b] = sb[col_slot<SWZ>(j + b * LR) * CW];
        }
    }
    if (MODE == 0) __syncthreads();  // in place: every input of the stage is in registers before the first output is written
    if (act) {
        int jhi = j, k = 0;
        if (Ns > 1) {
            jhi = j / Ns;
            k = j - jhi * Ns;
            const int q1 = k * (LR / Ns);  // W_L^(b k tstride), tstride = L / (Ns R)
            if (TWP && R == 16) {
                float2 w[16];
                w[1] = __ldg(tw + q1);
                w[4] = __ldg(tw + 4 * q1);
                w[2] = cmul(w[1], w[1]);
                w[3] = cmul(w[2], w[1]);
                w[8] = cmul(w[4], w[4]);
                w[12] = cmul(w[8], w[4]);
#pragma unroll
                for (int a = 4; a < 16; a += 4) {
#pragma unroll
                    for (int b = 1; b < 4; ++b) w[a + b] = cmul(w[a], w[b]);
                }
#pragma unroll
                for (int b = 1; b < R; ++b) v[b] = cmul(v[b], w[b < 16 ? b : 0]);
            } else if (TWP && R == 8) {
                float2 w[8];
                w[1] = __ldg(tw + q1);
                w[2] = cmul(w[1], w[1]);
                w[3] = cmul(w[2], w[1]);
                w[4] = cmul(w[2], w[2]);
                w[5] = cmul(w[4], w[1]);
                w[6] = cmul(w[4], w[2]);
                w[7] = cmul(w[4], w[3]);
#pragma unroll
                for (int b = 1; b < R; ++b) v[b] = cmul(v[b], w[b < 8 ? b : 0]);
            } else {
#pragma unroll
                for (int b = 1; b < R; ++b) v[b] = cmul(v[b], __ldg(tw + b * q1));
            }
        }
        fft::Dft<R>::run(v);
        const int orow = jhi * R * Ns + k;
        if (MODE == 2) {  // the inverse runs the forward FFT on re / im swapped data: swap back on the way out
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const int row = orow + t * Ns;
                oa[row * C] = v[t].y * scale;
                if (ob) ob[row * C] = v[t].x * scale;
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int t = 0; t < R; ++t) __stcs(reinterpret_cast<float2 *>(oa + (orow + t * Ns) * C), make_float2(v[t].y * scale, v[t].x * scale));
        } else {
#pragma unroll
            for (int t = 0; t < R; ++t) sb[col_slot<SWZ>(orow + t * Ns) * CW] = v[t];
        }
    }
    if (MODE != 2 && MODE != 4) __syncthreads();
}

// radix dispatch: folds to a single call when R_ is a compile-time constant (shape-specialised instantiations)
#define FD_COL_STAGE(MODE, R_, NS_)                                                                                           \
    switch (R_) {                                                                                                             \
        case 2: col_stage<2, MODE>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale); break;                               \
        case 4: col_stage<4, MODE>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale); break;                               \
        case 8: col_stage<8, MODE, false, TWP16>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale); break;                               \
        case 16: col_stage<16, MODE, false, TWP16>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale); break;                \
        default:                                                                                                              \
            if (GENERAL) {                                                                                                    \
                switch (R_) {                                                                                                 \
                    case 3: col_stage<3, MODE>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale); break;                   \
                    case 5: col_stage<5, MODE>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale); break;                   \
                    case 7: col_stage<7, MODE>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale); break;                   \
                    case 9: col_stage<9, MODE>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale); break;                   \
                    case 11: col_stage<11, MODE>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale); break;                 \
                    case 13: col_stage<13, MODE>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale); break;                 \
                    default: col_stage<17, MODE>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale); break;                 \
                }                                                                                                             \
            }                                                                                                                 \
            break;                                                                                                            \
    }

// Compile-time shape of a specialised instantiation (all zero: everything comes from the run-time plan).  With the shape known, every
// shared / global offset of a butterfly is an immediate of its load / store and the radix dispatch disappears — the generic
// instantiation spends more instructions on addresses than on arithmetic.
template <int L_, int C_, int CW_, int SP_, int R0_, int R1_, int R2_, int R3_>
struct ColShape {
    static constexpr int L = L_, C = C_, CW = CW_, SP = SP_, R0 = R0_, R1 = R1_, R2 = R2_, R3 = R3_;
    static constexpr int NS = (R0_ > 0) + (R1_ > 0) + (R2_ > 0) + (R3_ > 0);
    static constexpr int RMIN = R0_ == 0 ? 1 : (R3_ ? R3_ : R2_ ? R2_ : R1_ ? R1_ : R0_);  // radices are sorted, largest first
    static constexpr int JMAX = R0_ == 0 ? 0 : L_ / RMIN;
};
using ColShapeAny = ColShape<0, 0, 0, 0, 0, 0, 0, 0>;

template <bool GENERAL, int MAXT, int MINB, class SH>
__global__ void __launch_bounds__(MAXT, MINB) rfft_cols_kernel(const float *__restrict__ x, float *__restrict__ out, const float2 *__restrict__ tw,
                                                               const ColPlan pl, const int B, const float *__restrict__ mean,
                                                               const float *__restrict__ stdv, const int inverse) {
    extern __shared__ float2 csm[];
    constexpr bool TWP16 = true;  // radix-16 twiddles from two loads (col_stage)
    constexpr bool FIX = SH::L > 0;
    const int L = FIX ? SH::L : pl.L, C = FIX ? SH::C : pl.C, CW = FIX ? SH::CW : pl.CW, Jmax = FIX ? SH::JMAX : pl.Jmax, SP = FIX ? SH::SP : pl.SP;
    const int r0 = FIX ? SH::R0 : pl.radix[0], r1 = FIX ? SH::R1 : pl.radix[1], r2 = FIX ? SH::R2 : pl.radix[2], r3 = FIX ? SH::R3 : pl.radix[3];
    const int ns = FIX ? SH::NS : pl.n_stages;
    const int tid = threadIdx.x;
    const int t2 = tid / CW, cc = tid - t2 * CW;
    const int p = t2 / Jmax, j = t2 - p * Jmax;
    const int c = blockIdx.x * CW + cc;  // column groups on grid.x: the CTAs that share the rows of a series pair run together (L2 hits)
    const long long sa = 2ll * ((long long)blockIdx.y * SP + p);  // series 2s and 2s + 1 share a complex sequence
    const bool on = p < SP && c < C && sa < B;
    const bool has_b = sa + 1 < B;
    const size_t off = (size_t)(on ? sa : 0) * L * C + (on ? c : 0);
    const float *xa = x + off, *xb = has_b ? xa + (size_t)L * C : nullptr;
    float *oa = out + off, *ob = has_b ? oa + (size_t)L * C : nullptr;
    float2 *sb = csm + (p < SP ? p : 0) * L * CW + cc;
    const int n_real = L / 2 + 1;  // == ceil((L+1)/2), fourier.py:59
    const float scale = 1.0f / sqrtf((float)L);

    if (!inverse) {
        FD_COL_STAGE(1, r0, 1)
        if (ns > 1) FD_COL_STAGE(0, r1, r0)
        if (ns > 2) FD_COL_STAGE(0, r2, r0 * r1)
        if (ns > 3) FD_COL_STAGE(0, r3, r0 * r1 * r2)
        // unpack the spectra of the two real series and write the packed-real layout (fourier.py:21-40)
        if (on) {
#pragma unroll 4
            for (int k = j; k < n_real; k += Jmax) {
                const float2 zk = sb[k * CW], zn = sb[(k ? L - k : 0) * CW];
                // X_a = (Z[k] + conj(Z[L-k])) / 2 ; X_b = (Z[k] - conj(Z[L-k])) / (2i)
                const float hs = 0.5f * scale;
                const float ar = hs * (zk.x + zn.x), ai = hs * (zk.y - zn.y);
                const float br = hs * (zk.y + zn.y), bi = hs * (zn.x - zk.x);
                const bool has_im = !(k == 0 || 2 * k == L);
                oa[k * C] = ar;
                if (has_im) oa[(n_real + k - 1) * C] = ai;
                if (ob) {
                    ob[k * C] = br;
                    if (has_im) ob[(n_real + k - 1) * C] = bi;
                }
            }
        }
    } else {
        // rebuild the full spectrum of both series from the packed layout (fourier.py:59-76), de-standardised first (cmd/sample.py:76-78),
        // pack Z[k] = X_a[k] + i X_b[k] and store it with re / im SWAPPED (the inverse transform is the forward FFT of the swapped data)
        if (on) {
            const float *mu = mean ? mean + c : nullptr, *sd = mean ? stdv + c : nullptr;
            // every load of a batch of UNR rows is issued before the first use (specialised shapes: all rows of the thread at once)
            constexpr int UNR = FIX ? (SH::L / 2 + SH::JMAX) / (SH::JMAX > 0 ? SH::JMAX : 1) : 4;
            for (int k0 = j; k0 < n_real; k0 += UNR * Jmax) {
                float ra[UNR], ia[UNR], rb[UNR], ib[UNR], s_r[UNR], m_r[UNR], s_i[UNR], m_i[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int k = k0 + u * Jmax;
                    const bool ok = k < n_real, has_im = ok && !(k == 0 || 2 * k == L);
                    const int ir = k * C, ii = (n_real + k - 1) * C;
                    ra[u] = ok ? xa[ir] : 0.f;
                    ia[u] = has_im ? xa[ii] : 0.f;
                    rb[u] = (ok && xb) ? xb[ir] : 0.f;
                    ib[u] = (has_im && xb) ? xb[ii] : 0.f;
                    s_r[u] = (ok && mu) ? sd[ir] : 1.f;
                    m_r[u] = (ok && mu) ? mu[ir] : 0.f;
                    s_i[u] = (has_im && mu) ? sd[ii] : 1.f;
                    m_i[u] = (has_im && mu) ? mu[ii] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int k = k0 + u * Jmax;
                    if (k < n_real) {
                        const bool has_im = !(k == 0 || 2 * k == L);
                        float a_r = ra[u], a_i = ia[u], b_r = rb[u], b_i = ib[u];
                        if (mu) {  // same two roundings as x * std + mean
                            a_r = __fadd_rn(__fmul_rn(a_r, s_r[u]), m_r[u]);
                            if (xb) b_r = __fadd_rn(__fmul_rn(b_r, s_r[u]), m_r[u]);
                            if (has_im) {
                                a_i = __fadd_rn(__fmul_rn(a_i, s_i[u]), m_i[u]);
                                if (xb) b_i = __fadd_rn(__fmul_rn(b_i, s_i[u]), m_i[u]);
                            }
                        }
                        sb[k * CW] = make_float2(a_i + b_r, a_r - b_i);
                        if (has_im) sb[(L - k) * CW] = make_float2(b_r - a_i, a_r + b_i);  // mirror bin: conjugate spectra
                    }
                }
            }
        }
        __syncthreads();
        if (ns == 1) {
            FD_COL_STAGE(2, r0, 1)
        } else {
            FD_COL_STAGE(0, r0, 1)
            if (ns == 2) {
                FD_COL_STAGE(2, r1, r0)
            } else {
                FD_COL_STAGE(0, r1, r0)
                if (ns == 3) {
                    FD_COL_STAGE(2, r2, r0 * r1)
                } else {
                    FD_COL_STAGE(0, r2, r0 * r1)
                    FD_COL_STAGE(2, r3, r0 * r1 * r2)
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// Long series with many channels (cfg 5: max_len 4096, 16 channels): CHANNEL-pair packing, z = x[:, 2c] + i x[:, 2c+1] of one series.  A
// CTA transforms CW complex columns = 2 CW adjacent channels, so every row access of the CTA is 8 CW contiguous bytes — with CW = 4 a whole
// 32-byte sector.  (The series-pair packing above gives a CTA CW * 4 bytes of a row per series: at CW = 4 half of every sector is fetched for
// nothing, and ncu showed DRAM reads of 3x the algorithmic bytes at this shape — the L2 does not hold the rows until the other column groups
// come by.)  Same stages, same unpack algebra; only the global addressing differs.  Fixed shapes only (SH::L > 0), C even, C % (2 CW) == 0.
// ---------------------------------------------------------------------------------------------------------------------------------------
template <int MAXT, int MINB, class SH>
__global__ void __launch_bounds__(MAXT, MINB) rfft_cpair_kernel(const float *__restrict__ x, float *__restrict__ out, const float2 *__restrict__ tw,
                                                                const int B, const float *__restrict__ mean, const float *__restrict__ stdv,
                                                                const int inverse) {
    extern __shared__ float2 csm[];
    constexpr int L = SH::L, C = SH::C, CW = SH::CW, Jmax = SH::JMAX;
    constexpr int r0 = SH::R0, r1 = SH::R1, r2 = SH::R2, r3 = SH::R3, ns = SH::NS;
    static_assert(r0 == 16 && r1 == 16 && r2 == 16 && r3 == 0, "three radix-16 stages");
    constexpr bool SWZ = CW == 4;
    (void)ns;
#define FD_CP_STAGE(MODE, NS_) col_stage<16, MODE, SWZ, true>(sb, CW, L, NS_, j, on, tw, xa, xb, oa, ob, C, scale);
    const int tid = threadIdx.x;
    const int j = tid / CW, cc = tid - j * CW;
    const int c = 2 * (blockIdx.x * CW + cc);  // my complex column = channels c, c + 1
    const bool on = j < Jmax;
    const size_t off = (size_t)blockIdx.y * L * C + c;
    const float *xa = x + off, *xb = nullptr;
    float *oa = out + off, *ob = nullptr;
    float2 *sb = csm + cc;
    constexpr int n_real = L / 2 + 1;
    const float scale = 1.0f / sqrtf((float)L);
    (void)B;

    if (!inverse) {
        FD_CP_STAGE(3, 1)
        FD_CP_STAGE(0, 16)
        FD_CP_STAGE(0, 256)
        if (on) {  // X_a = (Z[k] + conj(Z[L-k])) / 2 -> channel c ; X_b = (Z[k] - conj(Z[L-k])) / (2i) -> channel c + 1 (fourier.py:21-40)
#pragma unroll 4
            for (int k = j; k < n_real; k += Jmax) {
                const float2 zk = sb[col_slot<SWZ>(k) * CW], zn = sb[col_slot<SWZ>(k ? L - k : 0) * CW];
                const float hs = 0.5f * scale;
                const float ar = hs * (zk.x + zn.x), ai = hs * (zk.y - zn.y);
                const float br = hs * (zk.y + zn.y), bi = hs * (zn.x - zk.x);
                __stcs(reinterpret_cast<float2 *>(oa + k * C), make_float2(ar, br));
                if (!(k == 0 || 2 * k == L)) __stcs(reinterpret_cast<float2 *>(oa + (n_real + k - 1) * C), make_float2(ai, bi));
            }
        }
    } else {
        if (on) {  // spectrum rebuild (fourier.py:59-76) after the de-standardisation (cmd/sample.py:76-78); stored re / im swapped
            const float *mu = mean ? mean + c : nullptr, *sd = mean ? stdv + c : nullptr;
            constexpr int UNR = 3;  // rows per batch: every load of a batch is issued before the first use (6 float2 per row with statistics)
#pragma unroll 1
            for (int j0 = j; j0 < n_real; j0 += UNR * Jmax) {
            float2 re[UNR], im[UNR], s_r[UNR], m_r[UNR], s_i[UNR], m_i[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int k = j0 + u * Jmax;
                const bool ok = k < n_real, has_im = ok && !(k == 0 || 2 * k == L);
                const int ir = k * C, ii = (n_real + k - 1) * C;
                const float2 zero = make_float2(0.f, 0.f), one = make_float2(1.f, 1.f);
                re[u] = ok ? __ldcs(reinterpret_cast<const float2 *>(xa + ir)) : zero;
                im[u] = has_im ? __ldcs(reinterpret_cast<const float2 *>(xa + ii)) : zero;
                s_r[u] = (ok && mu) ? __ldg(reinterpret_cast<const float2 *>(sd + ir)) : one;
                m_r[u] = (ok && mu) ? __ldg(reinterpret_cast<const float2 *>(mu + ir)) : zero;
                s_i[u] = (has_im && mu) ? __ldg(reinterpret_cast<const float2 *>(sd + ii)) : one;
                m_i[u] = (has_im && mu) ? __ldg(reinterpret_cast<const float2 *>(mu + ii)) : zero;
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int k = j0 + u * Jmax;
                if (k < n_real) {
                    const bool has_im = !(k == 0 || 2 * k == L);
                    float a_r = re[u].x, b_r = re[u].y, a_i = im[u].x, b_i = im[u].y;
                    if (mu) {  // same two roundings as x * std + mean
                        a_r = __fadd_rn(__fmul_rn(a_r, s_r[u].x), m_r[u].x);
                        b_r = __fadd_rn(__fmul_rn(b_r, s_r[u].y), m_r[u].y);
                        if (has_im) {
                            a_i = __fadd_rn(__fmul_rn(a_i, s_i[u].x), m_i[u].x);
                            b_i = __fadd_rn(__fmul_rn(b_i, s_i[u].y), m_i[u].y);
                        }
                    }
                    sb[col_slot<SWZ>(k) * CW] = make_float2(a_i + b_r, a_r - b_i);
                    if (has_im) sb[col_slot<SWZ>(L - k) * CW] = make_float2(b_r - a_i, a_r + b_i);
                }
            }
            }
        }
        __syncthreads();
        FD_CP_STAGE(0, 1)
        FD_CP_STAGE(0, 16)
        FD_CP_STAGE(4, 256)
#undef FD_CP_STAGE
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------------
// Lengths with a prime factor > 17 (US-Droughts' 365 = 5 x 73, prime 251, ...): Bluestein's chirp-z transform on top of the power-of-two
// stages above, all inside one CTA:  X[k] = w[k] . sum_n (z[n] w[n]) conj(w)[k - n],  w[n] = exp(-i pi n^2 / L)  (n^2 reduced mod 2L in
// integers, chirp and the spectrum H of the wrapped conj-chirp computed in fp64 on the host).  The length-M circular convolution
// (M = power of two >= 2L - 1) is two M-point FFTs in shared memory; the inverse FFT is the forward one on conjugated data.  Measured error
// is the same order as a plain fp32 FFT (2e-7 of the largest bin).  HBM traffic is unchanged: one coalesced pass in, one out.
// ---------------------------------------------------------------------------------------------------------------------------------------
template <int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 1024 ? 1 : MAXT == 512 ? 2 : 3) rfft_bluestein_kernel(const float *__restrict__ x, float *__restrict__ out, const float2 *__restrict__ tw,
                                                              const float2 *__restrict__ chirp, const float2 *__restrict__ Hf, const ColPlan pl,
                                                              const int B, const float *__restrict__ mean, const float *__restrict__ stdv,
                                                              const int inverse) {
    extern __shared__ float2 csm[];
    constexpr bool GENERAL = false, TWP16 = true;
    const int L = pl.L /* = M */, Ls = pl.Lseq, C = pl.C, CW = pl.CW, Jmax = pl.Jmax, SP = pl.SP;
    const int r0 = pl.radix[0], r1 = pl.radix[1], r2 = pl.radix[2], r3 = pl.radix[3], ns = pl.n_stages;
    const int tid = threadIdx.x;
    const int t2 = tid / CW, cc = tid - t2 * CW;
    const int p = t2 / Jmax, j = t2 - p * Jmax;
    const int c = blockIdx.x * CW + cc;
    const long long sa = 2ll * ((long long)blockIdx.y * SP + p);
    const bool on = p < SP && c < C && sa < B;
    const bool has_b = sa + 1 < B;
    const size_t off = (size_t)(on ? sa : 0) * Ls * C + (on ? c : 0);
    const float *xa = x + off, *xb = has_b ? xa + (size_t)Ls * C : nullptr;
    float *oa = out + off, *ob = has_b ? oa + (size_t)Ls * C : nullptr;
    float2 *sb = csm + (p < SP ? p : 0) * L * CW + cc;
    const int n_real = Ls / 2 + 1;
    const float scale = 1.0f / sqrtf((float)Ls);

    if (on) {
        if (!inverse) {
            for (int n0 = j; n0 < Ls; n0 += 4 * Jmax) {  // four rows of loads in flight
                float re[4], im[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int n = n0 + u * Jmax;
                    re[u] = n < Ls ? xa[n * C] : 0.f;
                    im[u] = (n < Ls && xb) ? xb[n * C] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int n = n0 + u * Jmax;
                    if (n < Ls) sb[n * CW] = cmul(make_float2(re[u], im[u]), __ldg(chirp + n));
                }
            }
        } else {
            const float *mu = mean ? mean + c : nullptr, *sd = mean ? stdv + c : nullptr;
            for (int k = j; k < n_real; k += Jmax) {  // spectrum rebuild as in rfft_cols_kernel (re / im swapped: inverse = forward of swapped data)
                const bool has_im = !(k == 0 || 2 * k == Ls);
                const int ir = k * C, ii = (n_real + k - 1) * C;
                float ra = xa[ir], ia = has_im ? xa[ii] : 0.f;
                float rb = xb ? xb[ir] : 0.f, ib = (xb && has_im) ? xb[ii] : 0.f;
                if (mu) {
                    const float s_r = sd[ir], m_r = mu[ir];
                    ra = __fadd_rn(__fmul_rn(ra, s_r), m_r);
                    if (xb) rb = __fadd_rn(__fmul_rn(rb, s_r), m_r);
                    if (has_im) {
                        const float s_i = sd[ii], m_i = mu[ii];
                        ia = __fadd_rn(__fmul_rn(ia, s_i), m_i);
                        if (xb) ib = __fadd_rn(__fmul_rn(ib, s_i), m_i);
                    }
                }
                sb[k * CW] = cmul(make_float2(ia + rb, ra - ib), __ldg(chirp + k));
                if (has_im) sb[(Ls - k) * CW] = cmul(make_float2(rb - ia, ra + ib), __ldg(chirp + Ls - k));
            }
        }
        for (int n = Ls + j; n < L; n += Jmax) sb[n * CW] = make_float2(0.f, 0.f);  // zero padding up to M
    }
    __syncthreads();
#define FD_BLUE_FFT                                \
    FD_COL_STAGE(0, r0, 1)                        \
    if (ns > 1) FD_COL_STAGE(0, r1, r0)           \
    if (ns > 2) FD_COL_STAGE(0, r2, r0 * r1)      \
    if (ns > 3) FD_COL_STAGE(0, r3, r0 * r1 * r2)
    FD_BLUE_FFT
    if (on)  // Y . H, conjugated for the inverse FFT (H carries the 1 / M)
        for (int k = j; k < L; k += Jmax) {
            const float2 v = cmul(sb[k * CW], __ldg(Hf + k));
            sb[k * CW] = make_float2(v.x, -v.y);
        }
    __syncthreads();
    FD_BLUE_FFT
#undef FD_BLUE_FFT
    // X[k] = conj(buf[k]) w[k]
    if (!inverse) {
        if (on)
            for (int k = j; k < Ls; k += Jmax) {
                const float2 v = sb[k * CW];
                sb[k * CW] = cmul(make_float2(v.x, -v.y), __ldg(chirp + k));
            }
        __syncthreads();
        if (on) {
            for (int k = j; k < n_real; k += Jmax) {  // unpack as in rfft_cols_kernel
                const float2 zk = sb[k * CW], zn = sb[(k ? Ls - k : 0) * CW];
                const float hs = 0.5f * scale;
                const bool has_im = !(k == 0 || 2 * k == Ls);
                oa[k * C] = hs * (zk.x + zn.x);
                if (has_im) oa[(n_real + k - 1) * C] = hs * (zk.y - zn.y);
                if (ob) {
                    ob[k * C] = hs * (zk.y + zn.y);
                    if (has_im) ob[(n_real + k - 1) * C] = hs * (zn.x - zk.x);
                }
            }
        }
    } else if (on) {
        for (int k = j; k < Ls; k += Jmax) {
            const float2 v = sb[k * CW];
            const float2 z = cmul(make_float2(v.x, -v.y), __ldg(chirp + k));
            oa[k * C] = z.y * scale;  // swap back
            if (ob) ob[k * C] = z.x * scale;
        }
    }
}

// max_len = R0 * R1 <= 32: thread = (series pair, column); the whole transform lives in registers
template <int R0, int R1>
__global__ void __launch_bounds__(128) rfft_small_kernel(const float *__restrict__ x, float *__restrict__ out, const float2 *__restrict__ tw,
                                                         const int B, const int C, const float *__restrict__ mean,
                                                         const float *__restrict__ stdv, const int inverse) {
    constexpr int L = R0 * R1, n_real = L / 2 + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pair = g / C;
    const int c = (int)(g - pair * C);
    const long long sa = 2 * pair;
    if (sa >= B) return;
    const bool has_b = sa + 1 < B;
    const size_t off = (size_t)sa * L * C + c;
    const float *xa = x + off, *xb = xa + (size_t)L * C;
    float *oa = out + off, *ob = oa + (size_t)L * C;
    const float scale = 1.0f / sqrtf((float)L);
    float2 z[L], w[L];
    if (!inverse) {
#pragma unroll
        for (int l = 0; l < L; ++l) {
            z[l].x = xa[l * C];
            z[l].y = has_b ? xb[l * C] : 0.f;
        }
    } else {
        const float *mu = mean ? mean + c : nullptr, *sd = mean ? stdv + c : nullptr;
#pragma unroll
        for (int k = 0; k < n_real; ++k) {
            const bool has_im = !(k == 0 || 2 * k == L);
            const int ir = k * C, ii = (n_real + k - 1) * C;
            float ra = xa[ir], ia = has_im ? xa[ii] : 0.f;
            float rb = has_b ? xb[ir] : 0.f, ib = (has_b && has_im) ? xb[ii] : 0.f;
            if (mu) {
                const float s_r = sd[ir], m_r = mu[ir];
                ra = ra * s_r + m_r;
                if (has_b) rb = rb * s_r + m_r;
                if (has_im) {
                    const float s_i = sd[ii], m_i = mu[ii];
                    ia = ia * s_i + m_i;
                    if (has_b) ib = ib * s_i + m_i;
                }
            }
            z[k] = make_float2(ia + rb, ra - ib);
            if (has_im) z[L - k] = make_float2(rb - ia, ra + ib);
        }
    }
    // stage 0: radix R0, inputs j + b R1 -> outputs j R0 + t (no twiddles)
#pragma unroll
    for (int j = 0; j < R1; ++j) {
        float2 v[R0];
#pragma unroll
        for (int b = 0; b < R0; ++b) v[b] = z[j + b * R1];
        fft::Dft<R0>::run(v);
#pragma unroll
        for (int t = 0; t < R0; ++t) w[j * R0 + t] = v[t];
    }
    // stage 1: radix R1 on W_L^(b j) w[j + b R0] -> outputs j + t R0
    if (R1 > 1) {
#pragma unroll
        for (int j = 0; j < R0; ++j) {
            float2 v[R1];
#pragma unroll
            for (int b = 0; b < R1; ++b) {
                v[b] = w[j + b * R0];
                if (b > 0 && j > 0) v[b] = cmul(v[b], __ldg(tw + b * j));
            }
            fft::Dft<R1>::run(v);
#pragma unroll
            for (int t = 0; t < R1; ++t) z[j + t * R0] = v[t];
        }
    } else {
#pragma unroll
        for (int l = 0; l < L; ++l) z[l] = w[l];
    }
    if (!inverse) {
#pragma unroll
        for (int k = 0; k < n_real; ++k) {
            const float2 zk = z[k], zn = z[k ? L - k : 0];
            const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
            const float br = 0.5f * (zk.y + zn.y), bi = -0.5f * (zk.x - zn.x);
            const bool has_im = !(k == 0 || 2 * k == L);
            oa[k * C] = ar * scale;
            if (has_im) oa[(n_real + k - 1) * C] = ai * scale;
            if (has_b) {
                ob[k * C] = br * scale;
                if (has_im) ob[(n_real + k - 1) * C] = bi * scale;
            }
        }
    } else {
#pragma unroll
        for (int l = 0; l < L; ++l) {
            oa[l * C] = z[l].y * scale;
            if (has_b) ob[l * C] = z[l].x * scale;
        }
    }
}

// radices of the column kernel for max_len L (largest first: the first stage has no twiddles); false: L has a prime factor > 17 or needs > 4 stages
static bool make_col_plan(int L, int C, ColPlan &pl, int max_threads = 512) {
    int n = L, n2 = 0;
    while (n % 2 == 0) { n /= 2; ++n2; }
    int rad[16], ns = 0;
    for (int f : {17, 13, 11, 9, 7, 5, 3})
        while (n % f == 0) {
            if (ns >= 8) return false;
            rad[ns++] = f;
            n /= f;
        }
    if (n != 1) return false;
    if (n2 > 0) {  // 2^n2 in ceil(n2 / 4) stages of near-equal size: 256 = 16 x 16, 512 = 8 x 8 x 8, 4096 = 16 x 16 x 16
        const int st = (n2 + 3) / 4, base = n2 / st, extra = n2 % st;
        for (int i = 0; i < st; ++i) {
            if (ns >= 8) return false;
            rad[ns++] = 1 << (base + (i < extra ? 1 : 0));
        }
    }
    if (ns == 0 || ns > 4) return false;
    std::sort(rad, rad + ns, [](int a, int b) { return a > b; });
    pl.n_stages = ns;
    int rmin = rad[0];
    for (int i = 0; i < ns; ++i) {
        pl.radix[i] = rad[i];
        rmin = std::min(rmin, rad[i]);
    }
    pl.L = L;
    pl.C = C;
    pl.Jmax = L / rmin;
    if (pl.Jmax > max_threads) return false;
    // columns per CTA: at most 512 threads and 64 KB per series pair, split evenly over the column groups
    int cw_max = std::min(max_threads / pl.Jmax, (int)(65536 / ((size_t)L * 8)));
    if (cw_max < 1) return false;
    const int ncg = (C + cw_max - 1) / cw_max;
    pl.CW = (C + ncg - 1) / ncg;
    // short series: several pairs per CTA (>= ~256 threads, <= 48 KB)
    pl.SP = 1;
    while (pl.CW == C && pl.SP < 16 && (pl.SP * 2) * pl.Jmax * pl.CW <= 256 && (size_t)(pl.SP * 2) * L * pl.CW * 8 <= 48 * 1024) pl.SP *= 2;
    return true;
}

template <bool GENERAL, int MAXT, int MINB, class SH>
static int launch_cols_inst(const float *x, float *out, const float2 *tw, const ColPlan &pl, int B, const float *mean, const float *stdv, bool inverse,
                            int dev, cudaStream_t s) {
    const int threads = ((pl.SP * pl.Jmax * pl.CW + 31) / 32) * 32;
    FD_CHECK(threads <= MAXT, "dft: plan of %d threads on a %d-thread kernel", threads, MAXT);
    const size_t smem = (size_t)pl.SP * pl.L * pl.CW * 8;
    const long long pairs = ((long long)B + 1) / 2;
    const long long gy = (pairs + pl.SP - 1) / pl.SP;
    FD_CHECK(gy <= 65535ll * 1024, "dft: batch too large");
    dim3 grid((unsigned)((pl.C + pl.CW - 1) / pl.CW), (unsigned)std::min<long long>(gy, 65535));
    static bool attr_set[64] = {false};  // (per instantiation)
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 64 && !attr_set[dev]) {
            FD_CUDA(cudaFuncSetAttribute(rfft_cols_kernel<GENERAL, MAXT, MINB, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[dev] = true;
        }
    }
    // grid.y is limited to 65535: larger batches go out in slices of series pairs
    for (long long y0 = 0; y0 < gy; y0 += 65535) {
        const long long ny = std::min<long long>(65535, gy - y0);
        grid.y = (unsigned)ny;
        const size_t skip = (size_t)y0 * pl.SP * 2 * pl.L * pl.C;
        rfft_cols_kernel<GENERAL, MAXT, MINB, SH><<<grid, threads, smem, s>>>(x + skip, out + skip, tw, pl, (int)(B - y0 * pl.SP * 2), mean, stdv, inverse ? 1 : 0);
    }
    return 0;
}

template <int MAXT, int MINB, class SH>
static int launch_cpair_inst(const float *x, float *out, const float2 *tw, int B, const float *mean, const float *stdv, bool inverse, int dev,
                             cudaStream_t s) {
    static_assert(SH::C % (2 * SH::CW) == 0 && SH::JMAX * SH::CW <= MAXT, "channel groups");
    constexpr size_t smem = (size_t)SH::L * SH::CW * 8;
    static bool attr_set[64] = {false};  // (per instantiation)
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 64 && !attr_set[dev]) {
            FD_CUDA(cudaFuncSetAttribute(rfft_cpair_kernel<MAXT, MINB, SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[dev] = true;
        }
    }
    for (long long y0 = 0; y0 < B; y0 += 65535) {  // grid.y is limited to 65535
        const int ny = (int)std::min<long long>(65535, B - y0);
        const size_t skip = (size_t)y0 * SH::L * SH::C;
        rfft_cpair_kernel<MAXT, MINB, SH><<<dim3(SH::C / (2 * SH::CW), ny), SH::JMAX * SH::CW, smem, s>>>(x + skip, out + skip, tw, ny, mean, stdv, inverse ? 1 : 0);
    }
    return 0;
}

static int launch_cols(const float *x, float *out, const float2 *tw, ColPlan pl, int B, const float *mean, const float *stdv, bool inverse, int dev,
                       cudaStream_t s) {
    // shape-specialised instantiations: the BASELINE configurations
    if (pl.L == 256 && pl.C == 12) {  // cfg 2: 16 x 16, one series pair (192 threads, 24 KB) per CTA
        using SH = ColShape<256, 12, 12, 1, 16, 16, 0, 0>;
        pl.CW = SH::CW, pl.SP = SH::SP;
        return launch_cols_inst<false, 192, 4, SH>(x, out, tw, pl, B, mean, stdv, inverse, dev, s);
    }
    if (pl.L == 4096 && pl.C == 16) {  // cfg 5: 16 x 16 x 16, four channel pairs (8 channels) of ONE series per CTA (1024 threads, 128 KB)
        // Measured at this shape (fraction of HBM peak, dft / idft): series-pair packing with 4 columns per CTA 0.34 / 0.30 (DRAM reads 3x the
        // algorithmic bytes: half sectors), with 2 columns and two CTAs per SM 0.24 / 0.21, DSMEM-cluster row scatter 0.21 / 0.14; channel-pair
        // packing 0.41 / 0.37, + swizzled rows and twiddle products 0.49 / 0.46 (ncu: DRAM bytes = algorithmic, L1 / shared pipe 50 %, issue
        // 41 %, one CTA per SM: the LDS / butterfly / STS phases of a stage do not overlap); 2 channel pairs per CTA and two CTAs per SM
        // 0.31 / 0.28; a persistent CTA that requests the next item's rows before the unpack phase 0.46 (the load latency is not the limiter).
        using SH = ColShape<4096, 16, 4, 1, 16, 16, 16, 0>;
        return launch_cpair_inst<1024, 1, SH>(x, out, tw, B, mean, stdv, inverse, dev, s);
    }
    if (pl.L == 187 && pl.C == 1) {  // the reference's ECG data set (MIT-BIH beats): 17 x 11, eight series pairs (136 threads) per CTA
        using SH = ColShape<187, 1, 1, 8, 17, 11, 0, 0>;
        pl.CW = SH::CW, pl.SP = SH::SP;
        return launch_cols_inst<true, 160, 4, SH>(x, out, tw, pl, B, mean, stdv, inverse, dev, s);
    }
    if (pl.L == 252 && pl.C == 5) {  // cfg 3: 9 x 7 x 4, one series pair (320 threads, 10 KB) per CTA
        using SH = ColShape<252, 5, 5, 1, 9, 7, 4, 0>;
        pl.CW = SH::CW, pl.SP = SH::SP;
        return launch_cols_inst<true, 320, 2, SH>(x, out, tw, pl, B, mean, stdv, inverse, dev, s);
    }
    bool pow2 = true;
    for (int i = 0; i < pl.n_stages; ++i) pow2 = pow2 && (pl.radix[i] & (pl.radix[i] - 1)) == 0;
    const int threads = pl.SP * pl.Jmax * pl.CW;
    if (pow2) {
        if (threads <= 256) return launch_cols_inst<false, 256, 3, ColShapeAny>(x, out, tw, pl, B, mean, stdv, inverse, dev, s);
        return launch_cols_inst<false, 512, 1, ColShapeAny>(x, out, tw, pl, B, mean, stdv, inverse, dev, s);
    }
    if (threads <= 256) return launch_cols_inst<true, 256, 2, ColShapeAny>(x, out, tw, pl, B, mean, stdv, inverse, dev, s);
    return launch_cols_inst<true, 512, 1, ColShapeAny>(x, out, tw, pl, B, mean, stdv, inverse, dev, s);
}

template <int R0, int R1>
static void launch_small(const float *x, float *out, const float2 *tw, int B, int C, const float *mean, const float *stdv, bool inverse, cudaStream_t s) {
    const long long items = (((long long)B + 1) / 2) * C;
    rfft_small_kernel<R0, R1><<<(unsigned)((items + 127) / 128), 128, 0, s>>>(x, out, tw, B, C, mean, stdv, inverse ? 1 : 0);
}

// ---- host side: plans and twiddle tables, cached per (device, L) ---------------------------------------------------------
struct FftCache {
    FftPlan plan;
    float2 *tw = nullptr;
    // Bluestein (lengths with a prime factor > 17): convolution length M, twiddles of M, chirp exp(-i pi n^2 / L), spectrum of the wrapped conj-chirp / M
    int blue_M = 0;
    float2 *blue_tw = nullptr, *blue_chirp = nullptr, *blue_H = nullptr;
};
static std::mutex g_fft_mu;
static std::map<std::pair<int, int>, FftCache> g_fft_cache;

static FftPlan make_plan(int L) {
    FftPlan p;
    p.n_stages = 0;
    int n = L;
    auto push = [&](int r) { p.radix[p.n_stages++] = r; };
    while (n % 8 == 0) { push(8); n /= 8; }
    while (n % 4 == 0) { push(4); n /= 4; }
    while (n % 2 == 0) { push(2); n /= 2; }
    for (int f = 3; (long long)f * f <= n; f += 2)
        while (n % f == 0) { push(f); n /= f; }
    if (n > 1) push(n);
    return p;
}

static void host_fft_pow2(std::vector<double> &re, std::vector<double> &im) {  // in-place radix-2 forward FFT, fp64 (table set-up only)
    const size_t n = re.size();
    for (size_t i = 1, jj = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; jj & bit; bit >>= 1) jj ^= bit;
        jj ^= bit;
        if (i < jj) {
            std::swap(re[i], re[jj]);
            std::swap(im[i], im[jj]);
        }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = -2.0 * M_PI / (double)len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const double wr = cos(ang * (double)k), wi = sin(ang * (double)k);
                const size_t a = i + k, b = i + k + len / 2;
                const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr;
                im[b] = im[a] - xi;
                re[a] += xr;
                im[a] += xi;
            }
    }
}

static int bluestein_setup(FftCache &c, int L) {
    int M = 1;
    while (M < 2 * L - 1) M <<= 1;
    std::vector<float2> tw(M), chirp(L), H(M);
    for (int q = 0; q < M; ++q) {
        const double a = -2.0 * M_PI * (double)q / (double)M;
        tw[q] = make_float2((float)cos(a), (float)sin(a));
    }
    std::vector<double> hr(M, 0.0), hi(M, 0.0);
    for (int n = 0; n < L; ++n) {
        const double ph = M_PI * (double)(((long long)n * n) % (2ll * L)) / (double)L;  // n^2 reduced mod 2L: exact phases
        chirp[n] = make_float2((float)cos(ph), (float)-sin(ph));
        hr[n] = cos(ph);
        hi[n] = sin(ph);
        if (n > 0) {
            hr[M - n] = cos(ph);
            hi[M - n] = sin(ph);
        }
    }
    host_fft_pow2(hr, hi);
    for (int k = 0; k < M; ++k) H[k] = make_float2((float)(hr[k] / M), (float)(hi[k] / M));
    FD_CUDA(cudaMalloc((void **)&c.blue_tw, M * sizeof(float2)));
    FD_CUDA(cudaMalloc((void **)&c.blue_chirp, L * sizeof(float2)));
    FD_CUDA(cudaMalloc((void **)&c.blue_H, M * sizeof(float2)));
    FD_CUDA(cudaMemcpy(c.blue_tw, tw.data(), M * sizeof(float2), cudaMemcpyHostToDevice));
    FD_CUDA(cudaMemcpy(c.blue_chirp, chirp.data(), L * sizeof(float2), cudaMemcpyHostToDevice));
    FD_CUDA(cudaMemcpy(c.blue_H, H.data(), M * sizeof(float2), cudaMemcpyHostToDevice));
    c.blue_M = M;
    return 0;
}

static int launch_bluestein(const FftCache &fc, const float *x, float *out, int B, int L, int C, const float *mean, const float *stdv, bool inverse,
                            int dev, cudaStream_t s) {
    ColPlan pl;
    FD_CHECK(make_col_plan(fc.blue_M, C, pl, 512), "dft: no plan for the Bluestein length %d", fc.blue_M);
    pl.Lseq = L;
    pl.SP = 1;
    while (pl.CW == C && pl.SP < 16 && (pl.SP * 2) * pl.Jmax * pl.CW <= 256 && (size_t)(pl.SP * 2) * pl.L * pl.CW * 8 <= 48 * 1024) pl.SP *= 2;
    const int threads = ((pl.SP * pl.Jmax * pl.CW + 31) / 32) * 32;
    const size_t smem = (size_t)pl.SP * pl.L * pl.CW * 8;
    const long long pairs = ((long long)B + 1) / 2;
    const long long gy = (pairs + pl.SP - 1) / pl.SP;
    static bool attr_set[64] = {false};
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 64 && !attr_set[dev]) {
            FD_CUDA(cudaFuncSetAttribute(rfft_bluestein_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            FD_CUDA(cudaFuncSetAttribute(rfft_bluestein_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            FD_CUDA(cudaFuncSetAttribute(rfft_bluestein_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            attr_set[dev] = true;
        }
    }
    dim3 grid((unsigned)((C + pl.CW - 1) / pl.CW), 1);
    for (long long y0 = 0; y0 < gy; y0 += 65535) {
        grid.y = (unsigned)std::min<long long>(65535, gy - y0);
        const size_t skip = (size_t)y0 * pl.SP * 2 * L * C;
        const int Brem = (int)(B - y0 * pl.SP * 2);
        if (threads <= 256)
            rfft_bluestein_kernel<256><<<grid, threads, smem, s>>>(x + skip, out + skip, fc.blue_tw, fc.blue_chirp, fc.blue_H, pl, Brem, mean, stdv, inverse ? 1 : 0);
        else if (threads <= 512)
            rfft_bluestein_kernel<512><<<grid, threads, smem, s>>>(x + skip, out + skip, fc.blue_tw, fc.blue_chirp, fc.blue_H, pl, Brem, mean, stdv, inverse ? 1 : 0);
        else
            rfft_bluestein_kernel<1024><<<grid, threads, smem, s>>>(x + skip, out + skip, fc.blue_tw, fc.blue_chirp, fc.blue_H, pl, Brem, mean, stdv, inverse ? 1 : 0);
    }
    return 0;
}

int launch_dft(const float *x, float *out, int B, int L, int C, const float *mean, const float *stdv, bool inverse, cudaStream_t s) {
    FD_CHECK(L <= 8192, "dft: max_len %d > 8192 is not supported", L);
    if (L == 1) {  // rfft of a length-1 series is the identity (ortho scale 1)
        // handled by the general kernel too (0 stages), fall through
    }
    int dev = 0;
    FD_CUDA(cudaGetDevice(&dev));
    FftCache *fc;
    {
        std::lock_guard<std::mutex> lk(g_fft_mu);
        auto key = std::make_pair(dev, L);
        auto it = g_fft_cache.find(key);
        if (it == g_fft_cache.end()) {
            FftCache c;
            c.plan = make_plan(L);
            FD_CHECK(c.plan.n_stages <= FFT_MAX_STAGES, "dft: too many stages");
            std::vector<float2> tw(L);
            for (int q = 0; q < L; ++q) {
                double a = -2.0 * M_PI * (double)q / (double)L;
                tw[q] = make_float2((float)cos(a), (float)sin(a));
            }
            FD_CUDA(cudaMalloc((void **)&c.tw, L * sizeof(float2)));
            FD_CUDA(cudaMemcpy(c.tw, tw.data(), L * sizeof(float2), cudaMemcpyHostToDevice));
            it = g_fft_cache.emplace(key, c).first;
        }
        fc = &it->second;
    }
    if ((size_t)L * C < (1u << 30) / 4) {  // fast paths index a series slab with 32-bit offsets
        bool done = true;
        switch (L) {
            case 8: launch_small<8, 1>(x, out, fc->tw, B, C, mean, stdv, inverse, s); break;
            case 12: launch_small<4, 3>(x, out, fc->tw, B, C, mean, stdv, inverse, s); break;
            case 16: launch_small<8, 2>(x, out, fc->tw, B, C, mean, stdv, inverse, s); break;
            case 20: launch_small<4, 5>(x, out, fc->tw, B, C, mean, stdv, inverse, s); break;
            case 24: launch_small<8, 3>(x, out, fc->tw, B, C, mean, stdv, inverse, s); break;
            case 28: launch_small<4, 7>(x, out, fc->tw, B, C, mean, stdv, inverse, s); break;
            case 32: launch_small<8, 4>(x, out, fc->tw, B, C, mean, stdv, inverse, s); break;
            default: done = false; break;
        }
        ColPlan pl;
        if (!done && L > 32 && make_col_plan(L, C, pl)) {
            FD_TRY(launch_cols(x, out, fc->tw, pl, B, mean, stdv, inverse, dev, s));
            done = true;
        }
        if (!done && L > 32 && 2 * L - 1 <= 8192) {  // a prime factor > 17: Bluestein on the power-of-two stages
            {
                std::lock_guard<std::mutex> lk(g_fft_mu);
                if (fc->blue_M == 0) FD_TRY(bluestein_setup(*fc, L));
            }
            FD_TRY(launch_bluestein(*fc, x, out, B, L, C, mean, stdv, inverse, dev, s));
            done = true;
        }
        if (done) {
            cudaError_t e = cudaGetLastError();
            FD_CHECK(e == cudaSuccess, "dft kernel launch failed: %s", cudaGetErrorString(e));
            g_global_launches += 1;
            return 0;
        }
    }
    // everything else (max_len <= 32 without a register kernel, prime-factor lengths > 4096): the generic shared-memory kernel
    const int Ptot = (C + 1) / 2;
    const size_t budget = 200 * 1024;
    int Pc = Ptot;
    while (Pc > 1 && ((size_t)L * 8 + 2 * (size_t)L * Pc * 8) > budget) Pc = (Pc + 1) / 2;
    // Long series whose channel pairs do not fit twice: an all-radix-8 plan can run IN PLACE (one buffer, inputs staged through registers),
    // which lets twice as many channel pairs share a CTA — whole 32-byte sectors of every row instead of half ones.
    int inplace = 0;
    if (Pc < Ptot) {
        bool all8 = fc->plan.n_stages > 0;
        for (int i = 0; i < fc->plan.n_stages; ++i) all8 = all8 && fc->plan.radix[i] == 8;
        int Pc1 = Ptot;
        while (Pc1 > 1 && ((size_t)L * 8 + (size_t)L * Pc1 * 8) > budget) Pc1 = (Pc1 + 1) / 2;
        if (all8 && Pc1 > Pc && (size_t)(L / 8) * Pc1 <= 4 * 512) {
            inplace = 1;
            Pc = Pc1;
        }
    }
    // short series: several series per CTA, so that a CTA works on ~4 K complex elements (and at most ~48 KB, four CTAs per SM)
    int S = 1;
    while (!inplace && S < 64 && S * 2 <= B && (size_t)L * Pc * (S * 2) <= 4096 && ((size_t)L * 8 + 2 * (size_t)L * Pc * (S * 2) * 8) <= 48 * 1024) S *= 2;
    size_t smem = (size_t)L * 8 + (inplace ? 1 : 2) * (size_t)L * Pc * S * 8 + 16;  // + the bulk-copy mbarrier
    FD_CHECK(smem <= budget + 16, "dft: max_len %d needs %zu bytes of shared memory", L, smem);
    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
        FD_CUDA(cudaFuncSetAttribute(rfft_packed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        FD_CUDA(cudaFuncSetAttribute(rfft_packed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        attr_set[dev] = true;
    }
    dim3 grid((B + S - 1) / S, (Ptot + Pc - 1) / Pc);
    // two radix-4 butterflies per thread and stage
    int threads = (int)(((size_t)L * Pc * S / 8 + 31) / 32) * 32;
    threads = threads < 64 ? 64 : threads > 512 ? 512 : threads;
    if (inplace)
        rfft_packed_kernel<true><<<grid, 512, smem, s>>>(x, out, fc->tw, fc->plan, B, L, C, Pc, S, mean, stdv, inverse ? 1 : 0);
    else
        rfft_packed_kernel<false><<<grid, threads, smem, s>>>(x, out, fc->tw, fc->plan, B, L, C, Pc, S, mean, stdv, inverse ? 1 : 0);
    cudaError_t e = cudaGetLastError();
    FD_CHECK(e == cudaSuccess, "dft kernel launch failed: %s", cudaGetErrorString(e));
    g_global_launches += 1;
    return 0;
}


// spectral_density of src/fdiff/utils/fourier.py:90-124 from the PACKED spectrum: out[b, k, c] = Re X_k^2 + Im X_k^2 for k = 0 .. L/2
// (Im X_0 = 0 and, for even L, Im X_{L/2} = 0 are not stored in the packed layout).  HBM-bound: reads 4 L C, writes 4 (L/2 + 1) C bytes per series.
__global__ void __launch_bounds__(256) spectral_density_kernel(const float *__restrict__ packed, float *__restrict__ out, int B, int L, int C) {
    const int n_real = L / 2 + 1, n_imag = (L - 1) / 2;
    const long long total = (long long)B * n_real * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long bk = i / C;
        const int k = (int)(bk % n_real), b = (int)(bk / n_real);
        const float *row = packed + (size_t)b * L * C;
        const float re = row[(size_t)k * C + c];
        const float im = (k >= 1 && k <= n_imag) ? row[(size_t)(n_real + k - 1) * C + c] : 0.f;
        out[i] = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));  // same two roundings as x_re**2 + x_im**2
    }
}

int launch_spectral_density(const float *packed, float *out, int B, int L, int C, cudaStream_t s) {
    const long long total = (long long)B * (L / 2 + 1) * C;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    spectral_density_kernel<<<grid, 256, 0, s>>>(packed, out, B, L, C);
    cudaError_t e = cudaGetLastError();
    FD_CHECK(e == cudaSuccess, "spectral_density_kernel launch failed: %s", cudaGetErrorString(e));
    g_global_launches += 1;
    return 0;
}

}  // namespace fd
