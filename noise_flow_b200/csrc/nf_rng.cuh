// Counter-based in-kernel normal sampler shared by the chain kernels (nf_kernels.cu, nf_wide.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nf {

// Philox4x32-10 (Salmon et al., SC'11); one call per pixel -> four normals by two Box-Muller pairs.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(unsigned int r) { return fmaf((float)r, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long offset,
                                                 unsigned long long patch, unsigned int pixel) {
    const uint4 r = philox4x32_10(make_uint4(pixel, (unsigned int)patch, (unsigned int)(patch >> 32), (unsigned int)offset),
                                  make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
    float4 o;
    float s, c;
    float rad = sqrtf(-2.f * __logf(u01(r.x)));
    sincospif(2.f * u01(r.y), &s, &c);
    o.x = rad * c; o.y = rad * s;
    rad = sqrtf(-2.f * __logf(u01(r.z)));
    sincospif(2.f * u01(r.w), &s, &c);
    o.z = rad * c; o.w = rad * s;
    return o;
}

}  // namespace nf
