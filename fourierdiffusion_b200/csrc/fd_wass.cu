// Sliced / marginal Wasserstein-2 distances between two sample sets — the metric stage that follows the sampler in cmd/sample.py:85
// (src/fdiff/sampling/metrics.py:102-199 on top of src/fdiff/utils/wasserstein.py:95-199, which calls POT's ot.emd2_1d per direction).
//
//   1. project: P[k][i] = <x_i, dir_k>  for K directions (fp32 samples, fp64 directions and accumulation like numpy's `data @ direction`),
//      or gather column k for the marginal distances (the directions are the standard basis, wasserstein.py:78-93);
//   2. sort every row of P (one row per direction): bitonic network, the sub-sequences that fit a CTA's shared memory in one kernel;
//   3. 1-D optimal transport between two sorted rows with uniform weights (emd2_1d, metric = squared Euclidean): the optimal plan is the
//      quantile coupling, so the cost is  sum over the merged quantile grid of  overlap x (a_i - b_j)^2 ; thread i owns the quantile
//      interval [i/n, (i+1)/n) of a and walks the b-intervals that intersect it (exact integer overlaps in units of 1/(n m)), fp64 sums.
// Everything is HBM / shared-memory bound integer-and-compare work; the projection is the only arithmetic and is tiny next to the sorts.
#include <math.h>
#include <stdint.h>

#include <algorithm>

#include "fd_common.cuh"

namespace fd {

// ---- 1. projections ---------------------------------------------------------------------------------------------------------------
// P[k][i] = sum_c x[i][c] * dir[k][c]   (x: (n, d) fp32 row-major, dir: (K, d) fp64).  Tile: 64 samples x 16 directions per CTA, the
// d dimension staged through shared memory 32 columns at a time; fp64 FMAs (the reference projects in float64).
__global__ void __launch_bounds__(256) wass_project_kernel(const float *__restrict__ x, const double *__restrict__ dir, float *__restrict__ P,
                                                           int n, int d, int K, long long ldp) {
    __shared__ float xs[64][33];
    __shared__ double ds[16][33];
    const int i0 = blockIdx.x * 64, k0 = blockIdx.y * 16;
    const int ti = threadIdx.x & 63, tk = threadIdx.x >> 6;  // thread: sample ti, directions tk, tk + 4, tk + 8, tk + 12
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c0 = 0; c0 < d; c0 += 32) {
        for (int e = threadIdx.x; e < 64 * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            xs[r][c] = (i0 + r < n && c0 + c < d) ? x[(size_t)(i0 + r) * d + c0 + c] : 0.f;
        }
        for (int e = threadIdx.x; e < 16 * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            ds[r][c] = (k0 + r < K && c0 + c < d) ? dir[(size_t)(k0 + r) * d + c0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {
            const double xv = (double)xs[ti][c];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fma(xv, ds[tk + 4 * q][c], acc[q]);
        }
        __syncthreads();
    }
    if (i0 + ti < n)
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (k0 + tk + 4 * q < K) P[(size_t)(k0 + tk + 4 * q) * ldp + i0 + ti] = (float)acc[q];
}

// marginal "projection": P[k][i] = x[i][k] (tiled transpose), rows padded to ldp
__global__ void __launch_bounds__(256) wass_transpose_kernel(const float *__restrict__ x, float *__restrict__ P, int n, int d, long long ldp) {
    __shared__ float t[32][33];
    const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int e = threadIdx.x; e < 1024; e += 256) {
        const int r = e >> 5, c = e & 31;
        t[r][c] = (i0 + r < n && k0 + c < d) ? x[(size_t)(i0 + r) * d + k0 + c] : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 1024; e += 256) {
        const int r = e >> 5, c = e & 31;  // r: feature, c: sample
        if (k0 + r < d && i0 + c < n) P[(size_t)(k0 + r) * ldp + i0 + c] = t[c][r];
    }
}

// rows [n, ldp) <- +inf so the padded power-of-two rows sort the real values to the front
__global__ void wass_pad_kernel(float *__restrict__ P, int n, long long ldp, int K) {
    const long long pad = ldp - n;
    const long long total = pad * K;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long k = e / pad, i = n + e % pad;
        P[k * ldp + i] = INFINITY;
    }
}

// ---- 2. bitonic sort of every row (length ldp = power of two) ------------------------------------------------------------------------
constexpr int SORT_TILE = 4096;  // elements of a row per CTA in the shared-memory kernel (16 KB)

__device__ __forceinline__ void cmp_swap(float &a, float &b, bool up) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = up ? lo : hi;
    b = up ? hi : lo;
}
// all network steps (k, j) with j < SORT_TILE for k in [k_first, k_last]: the tile stays in shared memory
__global__ void __launch_bounds__(1024) wass_sort_tile_kernel(float *__restrict__ P, long long ldp, int tile, int k_first, int k_last) {
    extern __shared__ float st[];
    float *row = P + (size_t)blockIdx.y * ldp + (size_t)blockIdx.x * tile;
    const int base = blockIdx.x * tile;
    for (int e = threadIdx.x; e < tile; e += blockDim.x) st[e] = row[e];
    __syncthreads();
    for (int k = k_first; k <= k_last; k <<= 1) {
        for (int j = min(k >> 1, tile >> 1); j > 0; j >>= 1) {
            for (int e = threadIdx.x; e < tile / 2; e += blockDim.x) {
                const int lo = 2 * e - (e & (j - 1));  // index with bit j clear
                const bool up = ((base + lo) & k) == 0;
                cmp_swap(st[lo], st[lo + j], up);
            }
            __syncthreads();
        }
    }
    for (int e = threadIdx.x; e < tile; e += blockDim.x) row[e] = st[e];
}
// one network step (k, j) with j >= SORT_TILE in global memory
__global__ void __launch_bounds__(256) wass_sort_global_kernel(float *__restrict__ P, long long ldp, int k, int j) {
    float *row = P + (size_t)blockIdx.y * ldp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ldp / 2; e += (long long)gridDim.x * blockDim.x) {
        const long long lo = 2 * e - (e & (j - 1));
        float a = row[lo], b = row[lo + j];
        cmp_swap(a, b, (lo & k) == 0);
        row[lo] = a;
        row[lo + j] = b;
    }
}

static int sort_rows(float *P, long long ldp, int K, cudaStream_t s) {
    const int tile = (int)std::min<long long>(ldp, SORT_TILE);
    const int threads = std::max(32, std::min(1024, tile / 2));
    dim3 gt((unsigned)(ldp / tile), (unsigned)K);
    // k = 2 .. tile: entirely inside a tile
    wass_sort_tile_kernel<<<gt, threads, tile * sizeof(float), s>>>(P, ldp, tile, 2, tile);
    g_global_launches += 1;
    for (long long k = 2ll * tile; k <= ldp; k <<= 1) {
        for (long long j = k >> 1; j >= tile; j >>= 1) {
            dim3 gg((unsigned)std::min<long long>((ldp / 2 + 255) / 256, 4096), (unsigned)K);
            wass_sort_global_kernel<<<gg, 256, 0, s>>>(P, ldp, (int)k, (int)j);
            g_global_launches += 1;
        }
        wass_sort_tile_kernel<<<gt, threads, tile * sizeof(float), s>>>(P, ldp, tile, (int)k, (int)k);
        g_global_launches += 1;
    }
    cudaError_t e = cudaGetLastError();
    FD_CHECK(e == cudaSuccess, "wasserstein sort launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// ---- 3. 1-D optimal transport between sorted rows, uniform weights (ot.emd2_1d, metric = 'sqeuclidean') --------------------------------
// out[k] = sqrt( sum_{i,j} |[i/n, (i+1)/n) ∩ [j/m, (j+1)/m)| (a_i - b_j)^2 ) / inv_scale[k]     (wasserstein.py:113-115, 139-141, 149-158)
__global__ void __launch_bounds__(256) wass_emd1d_kernel(const float *__restrict__ A, const float *__restrict__ Bm, long long lda, long long ldb,
                                                         int n, int m, const double *__restrict__ row_std, double *__restrict__ out) {
    const float *a = A + (size_t)blockIdx.x * lda, *b = Bm + (size_t)blockIdx.x * ldb;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double ai = (double)a[i];
        const long long lo = (long long)i * m, hi = (long long)(i + 1) * m;  // my interval in units of 1 / (n m)
        for (long long j = lo / n; j * n < hi; ++j) {
            const long long ov = min(hi, (j + 1) * (long long)n) - max(lo, j * (long long)n);
            const double dv = ai - (double)b[j];
            acc += (double)ov * dv * dv;
        }
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double w = sqrt(red[0] / ((double)n * (double)m));
        if (row_std) w /= row_std[blockIdx.x];  // 'standardise': both samples divided by the std of the original projection
        out[blockIdx.x] = w;
    }
}

// population standard deviation (np.std) of the first n entries of every row, fp64
__global__ void __launch_bounds__(256) wass_row_std_kernel(const float *__restrict__ A, long long lda, int n, double *__restrict__ out) {
    const float *a = A + (size_t)blockIdx.x * lda;
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)a[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
        __syncthreads();
    }
    const double mean = red[0] / n;
    __syncthreads();
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double dlt = (double)a[i] - mean;
        v += dlt * dlt;
    }
    red[threadIdx.x] = v;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sqrt(red[0] / n);
}

static long long pow2_ge(long long v) {
    long long p = 1;
    while (p < v) p <<= 1;
    return p;
}

// dirs_dev == nullptr: marginal distances (K must equal d).  work_dev: K * (pow2(n) + pow2(m)) floats + K doubles.
int launch_wasserstein(const float *x, const float *y, const double *dirs, int n, int m, int d, int K, int standardise, float *work,
                       double *out, cudaStream_t s) {
    const long long lda = pow2_ge(n), ldb = pow2_ge(m);
    float *A = work, *Bm = work + (size_t)K * lda;
    double *row_std = reinterpret_cast<double *>(work + (((size_t)K * (lda + ldb) + 1) & ~(size_t)1));  // 8-byte aligned
    if (dirs) {
        wass_project_kernel<<<dim3((n + 63) / 64, (K + 15) / 16), 256, 0, s>>>(x, dirs, A, n, d, K, lda);
        wass_project_kernel<<<dim3((m + 63) / 64, (K + 15) / 16), 256, 0, s>>>(y, dirs, Bm, m, d, K, ldb);
    } else {
        wass_transpose_kernel<<<dim3((n + 31) / 32, (d + 31) / 32), 256, 0, s>>>(x, A, n, d, lda);
        wass_transpose_kernel<<<dim3((m + 31) / 32, (d + 31) / 32), 256, 0, s>>>(y, Bm, m, d, ldb);
    }
    g_global_launches += 2;
    if (standardise) {  // np.std of the ORIGINAL projection (wasserstein.py:152-155), before the rows are sorted / padded
        wass_row_std_kernel<<<K, 256, 0, s>>>(A, lda, n, row_std);
        g_global_launches += 1;
    }
    if (lda > n) wass_pad_kernel<<<(unsigned)std::min<long long>(((lda - n) * K + 255) / 256, 8192), 256, 0, s>>>(A, n, lda, K);
    if (ldb > m) wass_pad_kernel<<<(unsigned)std::min<long long>(((ldb - m) * K + 255) / 256, 8192), 256, 0, s>>>(Bm, m, ldb, K);
    FD_TRY(sort_rows(A, lda, K, s));
    FD_TRY(sort_rows(Bm, ldb, K, s));
    wass_emd1d_kernel<<<K, 256, 0, s>>>(A, Bm, lda, ldb, n, m, standardise ? row_std : nullptr, out);
    g_global_launches += 1;
    cudaError_t e = cudaGetLastError();
    FD_CHECK(e == cudaSuccess, "wasserstein kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

size_t wasserstein_work_bytes(int n, int m, int K) {
    return ((size_t)K * (pow2_ge(n) + pow2_ge(m))) * sizeof(float) + (size_t)K * sizeof(double) + 16;
}

}  // namespace fd

// =======================================================================================================================================
// Data-set statistics of DiffusionDataset (src/fdiff/dataloaders/datamodules.py:42-65): per-feature mean and UNBIASED standard deviation
// over the series of a (n, L, C) tensor — the (L, C) statistics the sampler's de-standardisation consumes (cmd/sample.py:76-78) — and the
// standardisation (x - mean) / std itself.  HBM-bound: one pass over the data for the statistics (fp64 partial sums per slice of series,
// combined by a second tiny kernel), one for the standardisation.
// =======================================================================================================================================
namespace fd {

constexpr int STAT_SLICES = 64;

// partial[(slice * F + f) * 2 + {0, 1}] = sum, sum of squares of feature f over the series of the slice
__global__ void __launch_bounds__(256) feature_partial_kernel(const float *__restrict__ x, double *__restrict__ partial, long long n, int F) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const long long per = (n + gridDim.y - 1) / gridDim.y, i0 = per * blockIdx.y, i1 = min(n, i0 + per);
    double s = 0.0, q = 0.0;
    for (long long i = i0; i < i1; ++i) {
        const double v = (double)x[(size_t)i * F + f];
        s += v;
        q = fma(v, v, q);
    }
    partial[((size_t)blockIdx.y * F + f) * 2] = s;
    partial[((size_t)blockIdx.y * F + f) * 2 + 1] = q;
}
__global__ void __launch_bounds__(256) feature_finish_kernel(const double *__restrict__ partial, float *__restrict__ mean, float *__restrict__ stdv,
                                                             long long n, int F, int slices) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    double s = 0.0, q = 0.0;
    for (int k = 0; k < slices; ++k) {
        s += partial[((size_t)k * F + f) * 2];
        q += partial[((size_t)k * F + f) * 2 + 1];
    }
    const double m = s / (double)n;
    const double var = n > 1 ? fmax(q - (double)n * m * m, 0.0) / (double)(n - 1) : NAN;  // torch.std: Bessel's correction, NaN for one sample
    mean[f] = (float)m;
    stdv[f] = (float)sqrt(var);
}
// out = (x - mean) / std  (inverse = 0; datamodules.py:62)  or  x * std + mean  (inverse = 1; cmd/sample.py:76-78); same two roundings as torch
__global__ void __launch_bounds__(256) standardise_kernel(const float *__restrict__ x, const float *__restrict__ mean, const float *__restrict__ stdv,
                                                          float *__restrict__ out, long long total, int F, int inverse) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(e % F);
        out[e] = inverse ? __fadd_rn(__fmul_rn(x[e], stdv[f]), mean[f]) : __fdiv_rn(__fsub_rn(x[e], mean[f]), stdv[f]);
    }
}

int launch_feature_stats(const float *x, float *mean, float *stdv, long long n, int F, double *work, cudaStream_t s) {
    const int slices = (int)std::min<long long>(STAT_SLICES, std::max<long long>(1, n / 8));
    feature_partial_kernel<<<dim3((F + 255) / 256, slices), 256, 0, s>>>(x, work, n, F);
    feature_finish_kernel<<<(F + 255) / 256, 256, 0, s>>>(work, mean, stdv, n, F, slices);
    g_global_launches += 2;
    cudaError_t e = cudaGetLastError();
    FD_CHECK(e == cudaSuccess, "feature statistics launch failed: %s", cudaGetErrorString(e));
    return 0;
}
size_t feature_stats_work_bytes(int F) { return (size_t)STAT_SLICES * F * 2 * sizeof(double); }

int launch_standardise(const float *x, const float *mean, const float *stdv, float *out, long long n, int F, int inverse, cudaStream_t s) {
    const long long total = n * F;
    standardise_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 32), 256, 0, s>>>(x, mean, stdv, out, total, F, inverse);
    g_global_launches += 1;
    cudaError_t e = cudaGetLastError();
    FD_CHECK(e == cudaSuccess, "standardise launch failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace fd
