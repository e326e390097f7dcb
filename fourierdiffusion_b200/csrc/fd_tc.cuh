// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier, bulk async copy (TMA engine, UBLKCP),
// tcgen05 MMA / TMEM.  Thin inline-PTX wrappers; the encodings follow the PTX ISA (tcgen05 instruction descriptor,
// shared-memory matrix descriptor) — cross-checked against the bit-field layouts in CUTLASS' cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fd {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One arrival per WARP: every lane has issued its own fences (tcgen05.wait / tcgen05.fence::before_thread_sync) before the warp converges;
// the elected lane's arrive (release at CTA scope) then publishes the whole warp's writes.
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Post-mortem record of a protocol time-out: a pinned, mapped HOST buffer (it outlives the CUDA context that the trap below kills), one
// copy of the pointer per translation unit (set with cudaMemcpyToSymbol by the unit that wants records; nullptr = none).
// Layout (int32): [0] code (1 = mbarrier wait, 2 = dependency counter wait) [1] blockIdx.x [2] threadIdx.x [3] a [4] b [5] c [6] d.
static __device__ int *fd_abort_rec = nullptr;
// (NOT inlined: ~40 mbarrier wait sites per kernel would each carry ~90 cold instructions — half of the stack kernel's code — in between
//  the hot loops, and the instruction cache is a measured bottleneck of these kernels)
static __device__ __noinline__ void abort_with_record(int code, int a, int b, int c, int d) {
    volatile int *r = fd_abort_rec;
    if (r != nullptr && r[0] == 0) {  // (plain stores: the record lives in host memory; a lost race between two aborting threads is harmless)
        r[1] = (int)blockIdx.x;
        r[2] = (int)threadIdx.x;
        r[3] = a;
        r[4] = b;
        r[5] = c;
        r[6] = d;
        r[0] = code;
        __threadfence_system();
#pragma unroll 1
        for (int i = 0; i < 64; ++i) __nanosleep(1000);  // let the stores reach the host before the context goes down
    }
    __trap();
}
// Blocks until the phase with parity `parity` has completed.  A bounded spin turns a protocol bug into a trap (a CUDA
// error the host sees; the whole CUDA context of the process is lost) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    uint32_t spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > (1u << 24)) abort_with_record(1, (int)bar, (int)parity, 0, 0);
    }
}

// The same for a CONVERGED warp: the loop condition is a warp vote, so the compiler sees warp-uniform control flow behind the wait and keeps
// loop counters / operand addresses of the MMA issuer in uniform registers (no register -> uniform-register moves behind the barrier).
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (__all_sync(0xffffffffu, done != 0)) break;
        if (++spins > (1u << 24)) abort_with_record(1, (int)bar, (int)parity, 0, 0);
    }
}

// ---- bulk async copy global -> shared (TMA engine, no tensor map), completes on an mbarrier ---------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM allocation (one warp, .sync.aligned) -------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ---------------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"): the operand is stored as 16-byte k-chunks,
//   element (row, k) at  base + (k/4)*lbo + (row/8)*sbo + (row%8)*16 + (k%4)*4   (fp32/tf32: 4 elements per 16 B),
// i.e. core matrices of 8 rows x 16 B are 128 contiguous bytes; `lbo` = byte stride between the two k-chunks one MMA (K=8)
// consumes, `sbo` = byte stride between consecutive 8-row groups.  bits: [0,14) addr>>4, [16,30) lbo>>4, [32,46) sbo>>4,
// [46,48) version = 1 (Blackwell), [61,64) layout type = 0 (no swizzle).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major.
// bits: [4,6) D format (1 = F32), [7,10) A format (2 = TF32), [10,13) B format (2 = TF32), [15] A major, [16] B major (0 = K),
// [17,23) N>>3, [24,29) M>>4.
// high word of make_smem_desc (SBO + descriptor version)
__host__ __device__ constexpr uint32_t smem_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Instruction descriptor for kind::f16 with fp16 A and B (K-major), fp32 accumulate: A/B format 0 = F16.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA -----------------------------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T      (A: M x 8 K-major, B: N x 8 K-major, tf32)
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T      (A: 128 lanes x 8 columns of 32-bit tf32 values)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Predicated forms for warp-uniform issue loops: every lane computes the (uniform) operands, only the elected lane (`pred` != 0)
// issues — keeps descriptor arithmetic out of a divergent branch so it stays cheap.
__device__ __forceinline__ void mma_tf32_ss_if(uint32_t pred, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(pred)
        : "memory");
}
__device__ __forceinline__ void mma_tf32_ts_if(uint32_t pred, uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(pred)
        : "memory");
}
// kind::f16: A = 128 lanes x 8 TMEM columns holding 16 fp16 (two per 32-bit column, low half = even k), B = N x 16 fp16 K-major in smem
__device__ __forceinline__ void mma_f16_ts_if(uint32_t pred, uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(pred)
        : "memory");
}
// N MMAs of one GEMM (same D, same high descriptor word, per-k-step A address and low descriptor word) + the commit in ONE asm block:
// the D address, the high word and the instruction descriptor reach their uniform registers once instead of once per MMA.  The first MMA
// overwrites D when `acc_first` is 0.
template <uint32_t b_hi>
__device__ __forceinline__ void mma_f16_ts_x4_commit_if(uint32_t pred, uint32_t d_tmem, const uint32_t (&a)[4], const uint32_t (&lo)[4],
                                                        uint32_t idesc, uint32_t acc_first, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 b0, b1, b2, b3;\n\t"
        "setp.ne.b32 p, %11, 0;\n\t"
        "setp.ne.b32 q, %12, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "mov.b64 b0, {%5, %9};\n\t"
        "mov.b64 b1, {%6, %9};\n\t"
        "mov.b64 b2, {%7, %9};\n\t"
        "mov.b64 b3, {%8, %9};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], b0, %10, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], b1, %10, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%3], b2, %10, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%4], b3, %10, t;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%13];\n\t}" ::"r"(d_tmem),
        "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "n"(b_hi), "r"(idesc), "r"(acc_first), "r"(pred),
        "r"(bar)
        : "memory");
}
template <uint32_t b_hi>
__device__ __forceinline__ void mma_f16_ts_x5_commit_if(uint32_t pred, uint32_t d_tmem, const uint32_t (&a)[5], const uint32_t (&lo)[5],
                                                        uint32_t idesc, uint32_t acc_first, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 b0, b1, b2, b3, b4;\n\t"
        "setp.ne.b32 p, %13, 0;\n\t"
        "setp.ne.b32 q, %14, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "mov.b64 b0, {%6, %11};\n\t"
        "mov.b64 b1, {%7, %11};\n\t"
        "mov.b64 b2, {%8, %11};\n\t"
        "mov.b64 b3, {%9, %11};\n\t"
        "mov.b64 b4, {%10, %11};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], b0, %12, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], b1, %12, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%3], b2, %12, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%4], b3, %12, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%5], b4, %12, t;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%15];\n\t}" ::"r"(d_tmem),
        "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "n"(b_hi), "r"(idesc),
        "r"(acc_first), "r"(pred), "r"(bar)
        : "memory");
}
// four MMAs of one GEMM without the commit
template <uint32_t b_hi>
__device__ __forceinline__ void mma_f16_ts_x4_if(uint32_t pred, uint32_t d_tmem, const uint32_t (&a)[4], const uint32_t (&lo)[4], uint32_t idesc,
                                                 uint32_t acc_first) {
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 b0, b1, b2, b3;\n\t"
        "setp.ne.b32 p, %11, 0;\n\t"
        "setp.ne.b32 q, %12, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "mov.b64 b0, {%5, %9};\n\t"
        "mov.b64 b1, {%6, %9};\n\t"
        "mov.b64 b2, {%7, %9};\n\t"
        "mov.b64 b3, {%8, %9};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], b0, %10, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], b1, %10, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%3], b2, %10, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%4], b3, %10, t;\n\t}" ::"r"(d_tmem),
        "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "n"(b_hi), "r"(idesc), "r"(acc_first), "r"(pred)
        : "memory");
}
// nine kind::tf32 SS MMAs of one GEMM (K = 72): per-k-step low descriptor words a_lo0 + ks * a_step / b_lo0 + ks * b_step are formed inside the
// block from two registers each; `commit` != 0 appends the commit
template <uint32_t a_hi, uint32_t b_hi, uint32_t a_step, uint32_t b_step>
__device__ __forceinline__ void mma_tf32_ss_x9_if(uint32_t pred, uint32_t d_tmem, uint32_t a_lo0, uint32_t b_lo0, uint32_t idesc, uint32_t bar,
                                                  uint32_t commit) {
    asm volatile(
        "{\n\t.reg .pred p, q, t, c;\n\t.reg .b64 ad, bd;\n\t.reg .b32 al, bl;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "setp.ne.b32 q, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "setp.ne.and.b32 c, %6, 0, q;\n\t"
        "mov.b64 ad, {%1, %7};\n\tmov.b64 bd, {%2, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, p;\n\t"
        "add.u32 al, %1, %9;\n\tadd.u32 bl, %2, %10;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, t;\n\t"
        "add.u32 al, al, %9;\n\tadd.u32 bl, bl, %10;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, t;\n\t"
        "add.u32 al, al, %9;\n\tadd.u32 bl, bl, %10;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, t;\n\t"
        "add.u32 al, al, %9;\n\tadd.u32 bl, bl, %10;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, t;\n\t"
        "add.u32 al, al, %9;\n\tadd.u32 bl, bl, %10;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, t;\n\t"
        "add.u32 al, al, %9;\n\tadd.u32 bl, bl, %10;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, t;\n\t"
        "add.u32 al, al, %9;\n\tadd.u32 bl, bl, %10;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, t;\n\t"
        "add.u32 al, al, %9;\n\tadd.u32 bl, bl, %10;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, t;\n\t"
        "@c tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}" ::"r"(d_tmem),
        "r"(a_lo0), "r"(b_lo0), "r"(idesc), "r"(pred), "r"(bar), "r"(commit), "n"(a_hi), "n"(b_hi), "n"(a_step), "n"(b_step)
        : "memory");
}
// three kind::tf32 SS MMAs (consecutive k-steps) of one GEMM; the first accumulates when `acc_first` != 0; `commit` != 0 appends the commit
template <uint32_t a_hi, uint32_t b_hi, uint32_t a_step, uint32_t b_step>
__device__ __forceinline__ void mma_tf32_ss_x3_if(uint32_t pred, uint32_t d_tmem, uint32_t a_lo0, uint32_t b_lo0, uint32_t idesc, uint32_t acc_first,
                                                  uint32_t bar, uint32_t commit) {
    asm volatile(
        "{\n\t.reg .pred p, q, t, c;\n\t.reg .b64 ad, bd;\n\t.reg .b32 al, bl;\n\t"
        "setp.ne.b32 p, %11, 0;\n\t"
        "setp.ne.b32 q, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "setp.ne.and.b32 c, %6, 0, q;\n\t"
        "mov.b64 ad, {%1, %7};\n\tmov.b64 bd, {%2, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, p;\n\t"
        "add.u32 al, %1, %9;\n\tadd.u32 bl, %2, %10;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, t;\n\t"
        "add.u32 al, al, %9;\n\tadd.u32 bl, bl, %10;\n\tmov.b64 ad, {al, %7};\n\tmov.b64 bd, {bl, %8};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, t;\n\t"
        "@c tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}" ::"r"(d_tmem),
        "r"(a_lo0), "r"(b_lo0), "r"(idesc), "r"(pred), "r"(bar), "r"(commit), "n"(a_hi), "n"(b_hi), "n"(a_step), "n"(b_step), "r"(acc_first)
        : "memory");
}
// one kind::f16 MMA with both operands in shared memory (descriptors as low word + constant high word); `commit` != 0 appends the commit
template <uint32_t a_hi, uint32_t b_hi>
__device__ __forceinline__ void mma_f16_ss_lo_if(uint32_t pred, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate,
                                                 uint32_t bar, uint32_t commit) {
    asm volatile(
        "{\n\t.reg .pred p, q, c;\n\t.reg .b64 ad, bd;\n\t"
        "setp.ne.b32 p, %8, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "setp.ne.and.b32 c, %9, 0, q;\n\t"
        "mov.b64 ad, {%1, %3};\n\t"
        "mov.b64 bd, {%2, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n\t"
        "@c tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "n"(a_hi), "n"(b_hi), "r"(idesc), "r"(pred), "r"(bar), "r"(accumulate), "r"(commit)
        : "memory");
}
// one kind::tf32 MMA with both operands in shared memory (descriptors as low word + constant high word) and its commit
template <uint32_t a_hi, uint32_t b_hi>
__device__ __forceinline__ void mma_tf32_ss_commit_if(uint32_t pred, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 ad, bd;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "mov.b64 ad, {%1, %3};\n\t"
        "mov.b64 bd, {%2, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %5, p;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "n"(a_hi), "n"(b_hi), "r"(idesc), "r"(pred), "r"(bar)
        : "memory");
}
// keeps a value materialised in a register at this point of the program (operands prepared BEFORE a spin wait stay before it)
__device__ __forceinline__ void pin_reg(uint32_t &x) { asm volatile("" : "+r"(x)); }
// kind::f16, both operands from shared memory: A = M x 16 fp16 K-major, B = N x 16 fp16 K-major (two 16-byte k-chunks per MMA)
__device__ __forceinline__ void mma_f16_ss_if(uint32_t pred, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(pred)
        : "memory");
}
__device__ __forceinline__ void mma_commit_if(uint32_t pred, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar),
        "r"(pred)
        : "memory");
}
// mbarrier arrive once every tcgen05 operation issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- TMEM <-> registers: each thread of warp w reads lane 32*(w%4)+laneid, N consecutive 32-bit columns ---------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// wait for every outstanding tcgen05.ld of this thread; the registers of the load about to be consumed are operands, so that no use of
// them can be scheduled ahead of the wait (needed when another load is kept in flight across it)
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}

// ---- persistent-kernel plumbing: barrier recycling, cross-CTA dependency counters in global memory -----------------------------------
// mbarrier.inval before re-initialising a barrier word that held a (quiescent) barrier of the previous work item
__device__ __forceinline__ void mbar_reinit(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Blocks until *p >= target (counters only grow).  Bounded: a scheduling bug becomes a trap (CUDA error), not a hung GPU.
__device__ __forceinline__ void wait_counter_ge(const unsigned *p, unsigned target, int tag = 0) {
    unsigned spins = 0;
    unsigned v;
    while ((int)((v = ld_acquire_gpu(p)) - target) < 0) {
        __nanosleep(64);
        if (++spins > (1u << 23)) abort_with_record(2, tag, (int)v, (int)target, 0);
    }
}
// publish: every global write of this CTA that happened before (bar.sync-ordered) becomes visible to whoever acquires the counter
__device__ __forceinline__ void signal_counter(unsigned *p) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(1u) : "memory");
}
// generic-proxy <-> async-proxy ordering for GLOBAL data that another CTA produced with ordinary stores and this CTA stages with bulk copies
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// round-to-nearest (ties away) fp32 -> tf32, returned as fp32 bits with the low 13 mantissa bits cleared
__device__ __forceinline__ uint32_t f32_to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// cheap tf32 rounding for finite values: add half an ulp of the 10-bit mantissa to the magnitude (= round to nearest, ties away,
// exactly what cvt.rna.tf32.f32 does); the tensor core ignores the low 13 bits, so they need not be cleared.
__device__ __forceinline__ uint32_t tf32_round_bits(float x) { return __float_as_uint(x) + 0x1000u; }

// {hi, lo} fp32 -> packed fp16x2 (round to nearest even); 2^x on both halves with one MUFU operation
__device__ __forceinline__ uint32_t pack_f16x2(float hi, float lo) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// saturating variants (|x| > 65504 -> +-65504 instead of inf); RELU clamps negatives (and NaN) to +0 first
__device__ __forceinline__ uint32_t pack_f16x2_sat(float hi, float lo) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack_f16x2_relu_sat(float hi, float lo) {
    uint32_t r;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// fp32 -> fp16 with saturation to +-65504 (one instruction)
__device__ __forceinline__ __half f32_to_f16_sat(float x) {
    unsigned short r;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
    return __ushort_as_half(r);
}
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
    uint32_t y;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

// ---- packed fp32 pairs (FADD2 / FFMA2: two fp32 operations per issue slot) and the 3-input maximum (FMNMX3) -------------------------
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

}  // namespace tc
}  // namespace fd
