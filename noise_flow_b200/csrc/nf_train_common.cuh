// Device helpers shared by the two backward implementations: nf_train.cu (one warp per patch, host-synchronous
// orchestration; the exactness reference) and nf_trainer.cu (one CTA per patch, device-resident train step).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_params.h"
#include "nf_train.h"

namespace nf {

struct __align__(16) TrainSmem {
    float4 zp[NF_PIXELS];      // z' = z_in . A   (un-padded, [row*32 + col])
    float4 h2[34 * 34];        // padded h2 image (ring = 0)
    float4 g[34 * 34];         // padded gradient image (g_pre3 in B1, g_c1 in B3)
};

__device__ __forceinline__ float tw_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ void acc_out(double* dst, float v, int lane) {   // warp-reduce, one fp64 atomic
    v = tw_sum(v);
    if (lane == 0 && v != 0.f) atomicAdd(dst, (double)v);
}
__device__ __forceinline__ float4 mix_fwd(float4 v, const float* A) {   // out[o] = sum_i v[i] * A[i][o]
    float r[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) r[o] = v.x * A[0 * 4 + o] + v.y * A[1 * 4 + o] + v.z * A[2 * 4 + o] + v.w * A[3 * 4 + o];
    return make_float4(r[0], r[1], r[2], r[3]);
}
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// c1(r, c)[o] from the z' image (SAME zero padding)
template <class SM>
__device__ __forceinline__ void conv1_at(const NfTrainCoupling& P, const SM& S, int r, int c, float (&c1)[4]) {
#pragma unroll
    for (int o = 0; o < 4; ++o) c1[o] = P.b1[o];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int rr = r + dy - 1;
        if (rr < 0 || rr > 31) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int cc = c + dx - 1;
            if (cc < 0 || cc > 31) continue;
            const float4 z = S.zp[rr * 32 + cc];
#pragma unroll
            for (int o = 0; o < 4; ++o) c1[o] = fmaf(z.x, P.w1[dy][dx][0][o], fmaf(z.y, P.w1[dy][dx][1][o], c1[o]));
        }
    }
}

// fill zp (mixed input) for patch p; returns nothing
__device__ __forceinline__ void load_mixed(const NfTrainCoupling& P, TrainSmem& S, const float4* zin, int lane) {
    for (int r = 0; r < 32; ++r) {
        float4 z = zin[r * 32 + lane];
        if (P.has_mix) z = mix_fwd(z, &P.A[0][0]);
        S.zp[r * 32 + lane] = z;
    }
    __syncwarp();
}

// recompute h1 (post BN-1 + ReLU) and the normalised c2hat at pixel (r, lane)
template <class SM>
__device__ __forceinline__ void net_to_c2hat(const NfTrainCoupling& P, const SM& S, int r, int lane, float (&c1hat)[4],
                                             float (&h1)[4], float (&c2hat)[4]) {
    float c1[4];
    conv1_at(P, S, r, lane, c1);
#pragma unroll
    for (int o = 0; o < 4; ++o) { c1hat[o] = (c1[o] - P.m1[o]) * P.is1[o]; h1[o] = fmaxf(c1hat[o], 0.f); }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float c2 = P.b2[o];
#pragma unroll
        for (int i = 0; i < 4; ++i) c2 = fmaf(h1[i], P.w2[i][o], c2);
        c2hat[o] = (c2 - P.m2[o]) * P.is2[o];
    }
}

}  // namespace nf
