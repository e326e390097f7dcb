// Legacy warp-level MMA issue rate on B200: clocks per mma.sync per SM sub-partition, for the two shapes a register-resident attention
// would use (m16n8k8 tf32 for Q.K^T with dh = 6, m16n8k16 f16 for P.V) at 1, 2, 4 warps per sub-partition, 4 independent accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/hmma tools/ubench/hmma.cu && tools/ubench/hmma
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int KIND>
__global__ void k(float *out, long long *cyc, int iters) {
    float c[4][4] = {};
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float *out;
    long long *cyc, h[148];
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (int kind = 0; kind < 2; ++kind)
        for (int warps = 4; warps <= 16; warps *= 2) {
            for (int rep = 0; rep < 2; ++rep) {
                if (kind == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
                else k<1><<<148, warps * 32>>>(out, cyc, iters);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double per = (double)h[0] / (iters * 4.0 * (warps / 4));
            printf("%s  %2d warps/SM (%d per sub-partition): %.2f clocks per mma.sync per sub-partition (%lld cycles total)\n",
                   kind ? "m16n8k16 f16 " : "m16n8k8 tf32 ", warps, warps / 4, per, h[0]);
        }
    return 0;
}
