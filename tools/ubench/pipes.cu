// Micro-benchmark: issue throughput (warp instructions per clock per SM sub-partition) of the instructions the attention softmax uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes tools/ubench/pipes.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHAINS 8

template <int OP>
__global__ void k(float *out, long long *cyc, float seed) {
    float a[CHAINS];
    uint32_t u[CHAINS];
    uint64_t p[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; u[i] = __float_as_uint(a[i]) & 0x3bff3bffu; p[i] = ((uint64_t)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
    const float c1 = seed * 0.999f;
    uint64_t cc = ((uint64_t)__float_as_uint(c1) << 32) | __float_as_uint(c1);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
            if (OP == 2) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i + 1) % CHAINS]));
            if (OP == 3) asm volatile("sub.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
            if (OP == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(a[(i + 1) % CHAINS]));
            if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) % CHAINS]));
            if (OP == 6) asm volatile("sub.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
            if (OP == 7) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(c1));
            if (OP == 8) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(cc));
            if (OP == 9) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
            if (OP == 10) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]));
            if (OP == 11) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i] + __uint_as_float(u[i]) + (float)(p[i] & 0xffff);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int warps_per_sm) {
    float *out; long long *cyc;
    int blocks = 148, threads = warps_per_sm * 32;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    k<OP><<<blocks, threads>>>(out, cyc, 0.5f); cudaDeviceSynchronize();
    k<OP><<<blocks, threads>>>(out, cyc, 0.5f); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    double winstr_per_smsp = (double)ITERS * CHAINS * warps_per_sm / 4.0;
    printf("%-28s warps/SM %2d: %.3f warp-instr/clk/SMSP  (%.2f clk per warp-instr)\n", name, warps_per_sm, winstr_per_smsp / avg, avg / winstr_per_smsp);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {8, 16}) {
        run<0>("ex2.approx.ftz.f32", w);
        run<1>("ex2.approx.f16x2", w);
        run<11>("ex2.approx.ftz.bf16x2", w);
        run<2>("cvt.rn.f16x2.f32", w);
        run<3>("sub.f32x2", w);
        run<6>("sub.f32", w);
        run<4>("max.f32 (3-input)", w);
        run<5>("max.f32", w);
        run<7>("fma.rn.f32", w);
        run<8>("fma.rn.f32x2", w);
        run<9>("max.f16x2", w);
        run<10>("add.f16x2", w);
    }
    return 0;
}
