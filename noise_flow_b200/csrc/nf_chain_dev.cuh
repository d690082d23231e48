// Device-side building blocks shared by the fused chain kernels (nf_kernels.cu: all-fp32 warp-per-patch kernel;
// nf_hybrid.cu: the same pass with conv-3 on the tensor cores): where the resident patch lives (ZStore), packed fp32
// arithmetic, the transcendentals, and the layer passes that do not involve the coupling net.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_params.h"
#include "nf_kernels.h"

namespace nf {

// Where a warp keeps its resident patch z[32 rows][32 lanes] (float4 per pixel; lane = image column):
//  NF_Z_IN_TMEM = 1: in TENSOR MEMORY.  z is lane-private (a lane only ever touches its own column), which is
//      exactly TMEM's access model: warp w owns lanes 32*(w%4).., 128 columns (32 rows x 4 channels); one
//      tcgen05.ld/st .32x32b.x4 moves one image row for the whole warp.
//  NF_Z_IN_TMEM = 0: in shared memory (z[row*32 + lane]).
struct ZStore {
#if NF_Z_IN_TMEM
    uint32_t taddr;   // (first TMEM lane of this warp << 16) | first column of this warp's patch
    __device__ __forceinline__ void issue_ld(int r, uint32_t (&v)[4]) const {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr + (uint32_t)(r * 4)));
    }
    __device__ __forceinline__ float4 load(int r) const {
        uint32_t v[4];
        issue_ld(r, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        return make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
    }
    __device__ __forceinline__ void load2(int r1, int r2, float4& a, float4& b) const {
        uint32_t v[4], w[4];
        issue_ld(r1, v);
        issue_ld(r2, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        a = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
        b = make_float4(__uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]), __uint_as_float(w[3]));
    }
    __device__ __forceinline__ void store(int r, float4 z) const {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                     :: "r"(taddr + (uint32_t)(r * 4)), "r"(__float_as_uint(z.x)), "r"(__float_as_uint(z.y)),
                        "r"(__float_as_uint(z.z)), "r"(__float_as_uint(z.w)) : "memory");
    }
    __device__ __forceinline__ void commit() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#else
    float4* z;
    int lane;
    __device__ __forceinline__ float4 load(int r) const { return z[r * 32 + lane]; }
    __device__ __forceinline__ void load2(int r1, int r2, float4& a, float4& b) const { a = z[r1 * 32 + lane]; b = z[r2 * 32 + lane]; }
    __device__ __forceinline__ void store(int r, float4 v) const { z[r * 32 + lane] = v; }
    __device__ __forceinline__ void commit() const {}
#endif
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a);
    rb = *reinterpret_cast<unsigned long long*>(&b);
    rc = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }

// out[o] = sum_i v[i] * m[o][i]
__device__ __forceinline__ float4 mix4(float4 v, const float (&m)[4][4]) {
    const float2 lo = make_float2(v.x, v.y), hi = make_float2(v.z, v.w);
    float r[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float2 t = ffma2(lo, ld2(&m[o][0]), make_float2(0.f, 0.f));
        t = ffma2(hi, ld2(&m[o][2]), t);
        r[o] = t.x + t.y;
    }
    return make_float4(r[0], r[1], r[2], r[3]);
}

// tanh(v) = 1 - 2 / (exp(2v) + 1): ex2.approx + rcp.approx, ~1e-7 absolute error, saturates cleanly.
__device__ __forceinline__ float fast_tanh(float v) {
    const float e = exp2f(v * 2.885390081777927f);   // exp(2v); compiled with -use_fast_math -> ex2.approx
    return 1.f - __fdividef(2.f, e + 1.f);
}
__device__ __forceinline__ float fast_exp(float v) { return exp2f(v * 1.4426950408889634f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}


struct Acc4 { float2 v[4]; };   // one pending output row: 4 outputs x (even, odd) input-channel partial sums


// The 4x4 mix matrix as PER-THREAD registers.  Left to itself ptxas hoists these 16 uniform loads (and a dozen conv-1
// weights) out of the row loop, runs out of uniform registers, parks the values in vector registers and pays an R2UR per
// value and row step (31 of 293 instructions).  Adding a 0.0f that is only known at run time (the zero halo of the row
// ring, read from shared memory) makes the values genuinely per-thread for ptxas -- it sees through empty inline asm and
// even through shuffles of warp-uniform values -- so the mix runs as FFMA2 with vector-register operands: 261
// instructions per step, no R2UR (data -> latent kernel; the latent -> data kernel keeps its 34 R2UR either way).
template <bool INV, class CP>
__device__ __forceinline__ void load_mix_regs(const CP& P, float2 (&am)[4][2], const float rt_zero) {
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        am[o][0] = INV ? ld2(&P.a[o][0]) : ld2(&P.ainv[o][0]);
        am[o][1] = INV ? ld2(&P.a[o][2]) : ld2(&P.ainv[o][2]);
        am[o][0] = make_float2(am[o][0].x + rt_zero, am[o][0].y + rt_zero);
        am[o][1] = make_float2(am[o][1].x + rt_zero, am[o][1].y + rt_zero);
    }
}
__device__ __forceinline__ float4 mix4r(float4 v, const float2 (&m)[4][2]) {
    const float2 lo = make_float2(v.x, v.y), hi = make_float2(v.z, v.w);
    float r[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float2 t = ffma2(lo, m[o][0], make_float2(0.f, 0.f));
        t = ffma2(hi, m[o][1], t);
        r[o] = t.x + t.y;
    }
    return make_float4(r[0], r[1], r[2], r[3]);
}


// stand-alone 4x4 channel mix (Conv2d1x1 / tfb.Permute not followed by a coupling)
template <bool INV, class ZS>
__device__ __forceinline__ void mix_pass(const NfMixP& M, const ZS& zs) {
    float m[4][4];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int i = 0; i < 4; ++i) m[o][i] = INV ? M.a[o][i] : M.ainv[o][i];
    zs.commit();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) zs.store(r, mix4(zs.load(r), m));
    zs.commit();
}

// scale layers: sdn* (scale^2 = a*y + b) and gain* (scale = g)
template <bool INV, class ZS>
__device__ __forceinline__ void sdn_pass(const float4* __restrict__ yp, float a, float b, const ZS& zs, int lane, float& ldj) {
    float acc = 0.f;
    zs.commit();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
        const float4 y = __ldg(yp + r * 32 + lane);
        float4 z = zs.load(r);
        const float v0 = fmaf(a, y.x, b), v1 = fmaf(a, y.y, b), v2 = fmaf(a, y.z, b), v3 = fmaf(a, y.w, b);
        const float r0 = rsqrtf(v0), r1 = rsqrtf(v1), r2 = rsqrtf(v2), r3 = rsqrtf(v3);
        if (INV) { z.x *= r0; z.y *= r1; z.z *= r2; z.w *= r3; }                       // SdnEx5.py:125-126
        else     { z.x *= v0 * r0; z.y *= v1 * r1; z.z *= v2 * r2; z.w *= v3 * r3; }   // SdnEx5.py:106-107
        acc += (__logf(v0) + __logf(v1)) + (__logf(v2) + __logf(v3));
        zs.store(r, z);
    }
    zs.commit();
    ldj += INV ? -0.5f * acc : 0.5f * acc;                                              // SdnEx5.py:129 / :110
}

template <bool INV, class ZS>
__device__ __forceinline__ void gain_pass(float g, float ginv, float ldj_inv, const ZS& zs, int lane, float& ldj) {
    const float m = INV ? ginv : g;
    zs.commit();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
        float4 z = zs.load(r);
        z.x *= m; z.y *= m; z.z *= m; z.w *= m;
        zs.store(r, z);
    }
    zs.commit();
    if (lane == 0) ldj += INV ? ldj_inv : -ldj_inv;
}

}  // namespace nf
