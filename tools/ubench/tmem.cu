// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM for 1, 2, 4 warps per SM sub-partition (B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I fourierdiffusion_b200/csrc -o tools/ubench/tmem tools/ubench/tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fd_tc.cuh"
using namespace fd::tc;

#define ITERS 512
// MODE 0: ld x32 + wait each; 1: 4 lds then one wait; 2: st x32 + wait each; 3: ld x32, st x16 per iteration (softmax-like)
template <int MODE>
__global__ void k(float *out, long long *cyc) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t v[32], acc = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
    tmem_st32(tmem, v); tmem_st32(tmem + 32, v); tmem_st32(tmem + 64, v); tmem_st32(tmem + 96, v);
    tmem_st_wait();
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const uint32_t col = (it & 3) * 32 + (warp >> 2) * 128 % 512;
        if (MODE == 0) { tmem_ld32(tmem + col, v); tmem_ld_wait(); acc += v[0] + v[31]; }
        if (MODE == 1) {
            uint32_t a[32], b[32], c[32];
            tmem_ld32(tmem, v); tmem_ld32(tmem + 32, a); tmem_ld32(tmem + 64, b); tmem_ld32(tmem + 96, c); tmem_ld_wait();
            acc += v[0] + a[1] + b[2] + c[3];
        }
        if (MODE == 2) { v[0] = acc + it; tmem_st32(tmem + col, v); tmem_st_wait(); acc += 1; }
        if (MODE == 3) {
            tmem_ld32(tmem + col, v); tmem_ld_wait();
            uint32_t u[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = v[2 * i] ^ v[2 * i + 1];
            tmem_st16(tmem + col, u); tmem_st_wait(); acc += u[3];
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

template <int MODE>
void run(const char *name, int warps, double bytes_per_iter_per_warp) {
    float *out; long long *cyc;
    int blocks = 148, threads = warps * 32;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    k<MODE><<<blocks, threads>>>(out, cyc); cudaDeviceSynchronize();
    k<MODE><<<blocks, threads>>>(out, cyc); cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%-34s warps/SM %2d: %.1f cyc/iter/warp, %.1f B/cyc/SM  (%s)\n", name, warps, avg / ITERS, bytes_per_iter_per_warp * warps * ITERS / avg, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0>("ld.x32 + wait", w, 4096);
        run<1>("4 x ld.x32, one wait", w, 16384);
        run<2>("st.x32 + wait", w, 4096);
        run<3>("ld.x32, wait, st.x16, wait", w, 4096 + 2048);
    }
    return 0;
}
