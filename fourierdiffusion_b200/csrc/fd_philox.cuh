// Counter-based normal generator: Philox4x32-10 (Salmon et al., SC'11) + Box-Muller.
// Keyed by (seed) and counted by (group-of-4-elements, draw, global series index), so the noise a series sees does not
// depend on batch size or on how series are sharded over GPUs.  The reference draws from torch's global generator
// (sde.py:85,238); parity runs inject that noise instead, this generator is the performance-mode source.
#pragma once
#include <stdint.h>

namespace fd {

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(M0, c.x, hi0, lo0);
        philox_mulhilo(M1, c.z, hi1, lo1);
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// two uint32 -> two standard normals.  u1 in (0,1], u2 in [0,1).
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &z0, float &z1) {
    float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
    float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    float r = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    z0 = r * cs;
    z1 = r * sn;
}

// four standard normals for elements 4 group .. 4 group + 3 of `series` at draw `draw`
__device__ __forceinline__ void normals4(uint64_t seed, uint64_t series, uint32_t draw, uint32_t group, float z[4]) {
    uint4 r = philox4x32_10(make_uint4(group, draw, (uint32_t)series, (uint32_t)(series >> 32)),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    box_muller(r.x, r.y, z[0], z[1]);
    box_muller(r.z, r.w, z[2], z[3]);
}

}  // namespace fd
