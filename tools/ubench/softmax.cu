// Micro-benchmark of the softmax inner loops in isolation (no MMA, no hand-over): cycles per 32-key chunk per SM sub-partition for the
// row-maximum pass and the exponential pass at several polynomial fractions, with 1/2/4 warps per sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I fourierdiffusion_b200/csrc -I include -o tools/ubench/softmax tools/ubench/softmax.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fd_softmax.cuh"
using namespace fd;

#define ITERS 256
template <int MODE, int PN, int PD, int VAR>
__global__ void k(float *out, long long *cyc) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-0.37f * ((threadIdx.x * 7 + i * 13) % 29));
    for (int c = 0; c < 512; c += 32) tmem_st32(tmem + c, v);
    tmem_st_wait();
    __syncthreads();
    float m0 = -1e30f, m1 = m0, m2 = m0, m3 = m0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const int col = ((it + warp) & 7) * 64;
        if (MODE == 0) max_chunk<false>(tmem, col, 256, m0, m1, m2, m3);
        if (MODE == 1) { exp_chunk<false, PN, PD, VAR>(tmem, col, col + 32, 256, VAR == 3 ? 0.0f : 1.0f + (it & 1)); if (it & 1) tmem_st_wait(); }
    }
    tmem_st_wait();
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = m0 + m1 + m2 + m3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

template <int MODE, int PN, int PD, int VAR>
void run(const char *name, int warps) {
    float *out; long long *cyc;
    int blocks = 148, threads = warps * 32;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    k<MODE, PN, PD, VAR><<<blocks, threads>>>(out, cyc); cudaDeviceSynchronize();
    k<MODE, PN, PD, VAR><<<blocks, threads>>>(out, cyc); cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    double chunks_per_smsp = (double)ITERS * warps / 4.0;
    printf("%-26s warps/SM %2d: %7.1f cyc per chunk per warp, %6.1f cyc per chunk per SMSP = %.2f cyc/key  (%s)\n", name, warps, avg / ITERS,
           avg / chunks_per_smsp, avg / chunks_per_smsp / 32, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {8, 16}) {
        run<0, 0, 1, 0>("max pass", w);
        run<1, 5, 16, 1>("exp v1 poly 5/16", w);
        run<1, 3, 8, 1>("exp v1 poly 3/8", w);
        run<1, 2, 5, 1>("exp v1 poly 2/5", w);
        run<1, 0, 1, 3>("exp v3 poly 0", w);
        run<1, 1, 4, 3>("exp v3 poly 1/4", w);
        run<1, 5, 16, 3>("exp v3 poly 5/16", w);
        run<1, 3, 8, 3>("exp v3 poly 3/8", w);
        run<1, 2, 5, 3>("exp v3 poly 2/5", w);
        run<1, 7, 16, 3>("exp v3 poly 7/16", w);
        run<1, 1, 2, 3>("exp v3 poly 1/2", w);
    }
    return 0;
}
