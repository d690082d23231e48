// Device code of the CTA-per-patch chain kernels (nf_wide.cu: plain couplings; nf_wide_cond.cu: programs with
// clean-image-conditioned couplings).  See nf_wide.cu for the design notes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_kernels.h"
#include "nf_params.h"
#include "nf_rng.cuh"
#include "nf_wide.h"

namespace nf {

#define WIDE_THREADS 512
#define WIDE_WARPS (WIDE_THREADS / 32)

__device__ __forceinline__ float w_tanh(float v) {   // as nf_kernels.cu fast_tanh: ex2.approx + rcp.approx, ~1e-7 abs
    const float e = exp2f(v * 2.885390081777927f);
    return 1.f - __fdividef(2.f, e + 1.f);
}
__device__ __forceinline__ float w_exp(float v) { return exp2f(v * 1.4426950408889634f); }
__device__ __forceinline__ float w_warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
// out[o] = sum_i v[i] * m[o*4 + i]
__device__ __forceinline__ float4 w_mix(float4 v, const float* m) {
    float r[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) r[o] = v.x * m[o * 4] + v.y * m[o * 4 + 1] + v.z * m[o * 4 + 2] + v.w * m[o * 4 + 3];
    return make_float4(r[0], r[1], r[2], r[3]);
}

template <int W>
struct WideSmem {
    static constexpr int WMAX = NfWideLayoutG(W, 6, 4).size() > NfWideLayoutG(W, 4, 8).size() ? NfWideLayoutG(W, 6, 4).size()
                                                                                              : NfWideLayoutG(W, 4, 8).size();
    float4 z[NF_PIXELS];
    float4 y[NF_PIXELS];                 // clean patch, staged only for programs with clean-image-conditioned couplings
    float4 h2[34 * (W / 4) * 34];
    float w[WMAX > 128 ? WMAX : 128];
    float red[WIDE_WARPS * 4];
    float sacc[2 * W];
};

// deterministic CTA sum: fixed shuffle tree per warp, warps added in order; valid in thread 0
__device__ __forceinline__ float w_cta_sum(float v, float* red, int warp, int lane) {
    v = w_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float tot = 0.f;
    if (threadIdx.x == 0)
        for (int k = 0; k < WIDE_WARPS; ++k) tot += red[k];
    return tot;
}

// STAGE 0: the coupling.  STAGE 1 / 2 (batch-statistics probes): accumulate per-channel sum and sum of squares of
// the conv-1 / conv-2 output before BatchNorm into S.sacc and leave z untouched.
// MODE (NF_COUPLING_MODE_*): 0 = AffineCoupling, net(x0) transforms x1; 1 = AffineCouplingCondXY[G], net(concat(x0, yy))
// transforms x1; 2 = AffineCouplingCondY[G], net(yy) shifts / scales all four channels.
template <int W, int STAGE, int MODE>
__device__ __forceinline__ void wide_coupling(const float* __restrict__ gblob, WideSmem<W>& S, int warp, int lane, const bool INV,
                                              float& ldj) {
    constexpr int CIN = MODE == 0 ? 2 : (MODE == 1 ? 6 : 4), COUT = MODE == 2 ? 8 : 4;
    constexpr NfWideLayoutG L(W, CIN, COUT);
    constexpr int G = W / 4;
    for (int k = threadIdx.x * 4; k < L.size(); k += WIDE_THREADS * 4)
        *reinterpret_cast<float4*>(&S.w[k]) = *reinterpret_cast<const float4*>(&gblob[k]);
    __syncthreads();
    const float* w = S.w;
    const bool has_mix = w[L.META] != 0.f;
    if (INV && has_mix) {
        for (int r = warp; r < 32; r += WIDE_WARPS) S.z[r * 32 + lane] = w_mix(S.z[r * 32 + lane], w + L.A);
        __syncthreads();
    }
    // Each thread owns TWO pixels (rows `warp` and `warp + 16`, column `lane`) and walks them together: every weight
    // fetched from shared memory (broadcast LDS) feeds both, which halves the shared-memory instructions per FMA
    // (with one pixel the LDS pipe, 1 per clock and SM, saturates together with the FMA pipe).
    static_assert(WIDE_WARPS == 16, "two rows per warp");
    const int rows2[2] = {warp, warp + 16};
    // ---- P1: conv 3x3 SAME (CIN -> W) + folded BN + ReLU ; conv 1x1 (W -> W) + folded BN + ReLU
    {
        float h1[2][W];
#pragma unroll
        for (int o = 0; o < W; ++o) h1[0][o] = h1[1][o] = w[L.B1 + o];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int cc = lane + dx - 1;
                float in[2][CIN];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int rr = rows2[q] + dy - 1;
#pragma unroll
                    for (int ci = 0; ci < CIN; ++ci) in[q][ci] = 0.f;
                    if (rr >= 0 && rr <= 31 && cc >= 0 && cc <= 31) {
                        if (MODE != 2) { const float4 v = S.z[rr * 32 + cc]; in[q][0] = v.x; in[q][1] = v.y; }
                        if (MODE != 0) {
                            const float4 yv = S.y[rr * 32 + cc];
                            in[q][CIN - 4] = yv.x; in[q][CIN - 3] = yv.y; in[q][CIN - 2] = yv.z; in[q][CIN - 1] = yv.w;
                        }
                    }
                }
                const float* wt = w + L.w1() + (dy * 3 + dx) * W * CIN;
#pragma unroll
                for (int o = 0; o < W; ++o) {
#pragma unroll
                    for (int ci = 0; ci < CIN; ci += 2) {
                        const float2 wv = *reinterpret_cast<const float2*>(wt + o * CIN + ci);
#pragma unroll
                        for (int q = 0; q < 2; ++q) h1[q][o] = fmaf(in[q][ci], wv.x, fmaf(in[q][ci + 1], wv.y, h1[q][o]));
                    }
                }
            }
        }
        if (STAGE == 1) {   // probes reduce right away (no per-thread accumulator arrays: registers)
#pragma unroll
            for (int o = 0; o < W; ++o) {
                const float a = w_warp_sum(h1[0][o] + h1[1][o]), q = w_warp_sum(fmaf(h1[0][o], h1[0][o], h1[1][o] * h1[1][o]));
                if (lane == 0) { atomicAdd(&S.sacc[o], a); atomicAdd(&S.sacc[W + o], q); }
            }
            return;
        }
#pragma unroll
        for (int o = 0; o < W; ++o) { h1[0][o] = fmaxf(h1[0][o], 0.f); h1[1][o] = fmaxf(h1[1][o], 0.f); }
#pragma unroll
        for (int o = 0; o < W; o += 4) {
            float acc[2][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = w[L.b2() + o + j];
#pragma unroll
            for (int i = 0; i < W; i += 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 wv = *reinterpret_cast<const float4*>(w + L.w2() + (o + j) * W + i);
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        acc[q][j] = fmaf(h1[q][i], wv.x, fmaf(h1[q][i + 1], wv.y, fmaf(h1[q][i + 2], wv.z, fmaf(h1[q][i + 3], wv.w, acc[q][j]))));
                }
            }
            if (STAGE == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float a = w_warp_sum(acc[0][j] + acc[1][j]), q = w_warp_sum(fmaf(acc[0][j], acc[0][j], acc[1][j] * acc[1][j]));
                    if (lane == 0) { atomicAdd(&S.sacc[o + j], a); atomicAdd(&S.sacc[W + o + j], q); }
                }
            } else {
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    S.h2[((rows2[q] + 1) * G + (o >> 2)) * 34 + lane + 1] =
                        make_float4(fmaxf(acc[q][0], 0.f), fmaxf(acc[q][1], 0.f), fmaxf(acc[q][2], 0.f), fmaxf(acc[q][3], 0.f));
            }
        }
    }
    if (STAGE) return;
    __syncthreads();
    // ---- P2: conv 3x3 over the zero-padded h2 (+ folded edge-indicator bias) -> shift, log-scale ; affine update
    {
        const int cc = lane == 0 ? 0 : (lane == 31 ? 2 : 1);
        float pre[2][COUT];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int rc = rows2[q] == 0 ? 0 : (rows2[q] == 31 ? 2 : 1);
#pragma unroll
            for (int o = 0; o < COUT; ++o) pre[q][o] = w[L.B3 + (rc * 3 + cc) * COUT + o];
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float* wt = w + L.w3() + (dy * 3 + dx) * W * COUT;
#pragma unroll 4
                for (int g = 0; g < G; ++g) {
                    float4 h[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) h[q] = S.h2[((rows2[q] + dy) * G + g) * 34 + lane + dx];
#pragma unroll
                    for (int ob = 0; ob < COUT; ob += 4) {
                        const float4 w0 = *reinterpret_cast<const float4*>(wt + (g * 4 + 0) * COUT + ob);
                        const float4 w1 = *reinterpret_cast<const float4*>(wt + (g * 4 + 1) * COUT + ob);
                        const float4 w2 = *reinterpret_cast<const float4*>(wt + (g * 4 + 2) * COUT + ob);
                        const float4 w3 = *reinterpret_cast<const float4*>(wt + (g * 4 + 3) * COUT + ob);
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            pre[q][ob + 0] = fmaf(h[q].x, w0.x, fmaf(h[q].y, w1.x, fmaf(h[q].z, w2.x, fmaf(h[q].w, w3.x, pre[q][ob + 0]))));
                            pre[q][ob + 1] = fmaf(h[q].x, w0.y, fmaf(h[q].y, w1.y, fmaf(h[q].z, w2.y, fmaf(h[q].w, w3.y, pre[q][ob + 1]))));
                            pre[q][ob + 2] = fmaf(h[q].x, w0.z, fmaf(h[q].y, w1.z, fmaf(h[q].z, w2.z, fmaf(h[q].w, w3.z, pre[q][ob + 2]))));
                            pre[q][ob + 3] = fmaf(h[q].x, w0.w, fmaf(h[q].y, w1.w, fmaf(h[q].z, w2.w, fmaf(h[q].w, w3.w, pre[q][ob + 3]))));
                        }
                    }
                }
            }
        const float scale = w[L.META + 1];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float4 z = S.z[rows2[q] * 32 + lane];
            if (MODE == 2) {   // AffineCouplingCondY.py:44-72: every channel, shift = pre[0..3], log-scale = pre[4..7]
                float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float ls = scale * w_tanh(pre[q][4 + k]);
                    if (INV) { zz[k] = fmaf(zz[k], w_exp(ls), pre[q][k]); ldj += ls; }
                    else     { zz[k] = (zz[k] - pre[q][k]) * w_exp(-ls); ldj -= ls; }
                }
                z = make_float4(zz[0], zz[1], zz[2], zz[3]);
            } else {
                const float ls0 = scale * w_tanh(pre[q][2]), ls1 = scale * w_tanh(pre[q][3]);      // layers.py:362-365
                if (INV) {                                                                          // layers.py:355-375
                    z.z = fmaf(z.z, w_exp(ls0), pre[q][0]);
                    z.w = fmaf(z.w, w_exp(ls1), pre[q][1]);
                    ldj += ls0 + ls1;
                } else {                                                                            // layers.py:333-353
                    z.z = (z.z - pre[q][0]) * w_exp(-ls0);
                    z.w = (z.w - pre[q][1]) * w_exp(-ls1);
                    ldj -= ls0 + ls1;
                }
            }
            if (!INV && has_mix) z = w_mix(z, w + L.AINV);
            S.z[rows2[q] * 32 + lane] = z;
        }
    }
}

template <int W, int STAGE, bool COND>
__device__ __forceinline__ void wide_coupling_any(const float* __restrict__ pb, WideSmem<W>& S, int warp, int lane, bool inv, float& ldj) {
    if (COND) {
        const int mode = (int)__ldg(pb + NfWideLayoutG::META + 2);
        if (mode == NF_COUPLING_MODE_XY) { wide_coupling<W, STAGE, 1>(pb, S, warp, lane, inv, ldj); return; }
        if (mode == NF_COUPLING_MODE_Y) { wide_coupling<W, STAGE, 2>(pb, S, warp, lane, inv, ldj); return; }
    }
    wide_coupling<W, STAGE, 0>(pb, S, warp, lane, inv, ldj);
}

// The chain over one patch per CTA.  COND = false: plain AffineCoupling programs only (INV is a compile-time constant of the
// calling kernel); COND = true: programs with clean-image-conditioned couplings (nf_wide_cond.cu; direction at run time).
template <int W, bool COND>
__device__ __forceinline__ void wide_chain_body(const NfWideProgram& prog, const float* __restrict__ blob, const NfChainArgs& a, const bool INV) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WideSmem<W>& S = *reinterpret_cast<WideSmem<W>*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int G = W / 4;
    for (int k = threadIdx.x; k < 34 * G * 34; k += WIDE_THREADS) {   // zero padding ring of h2 (never written again)
        const int R = k / (G * 34), C = k % 34;
        if (R == 0 || R == 33 || C == 0 || C == 33) S.h2[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int k = threadIdx.x; k < 2 * W; k += WIDE_THREADS) S.sacc[k] = 0.f;
    __syncthreads();
    for (long long p = blockIdx.x; p < a.n; p += gridDim.x) {
        int row = a.rows ? a.rows[p] : a.default_row;
        row = min(max(row, 0), NF_MAX_ROWS - 1);
        for (int k = threadIdx.x; k < NF_PIXELS; k += WIDE_THREADS) {
            float4 v;
            if (a.in) {
                v = __ldcs(reinterpret_cast<const float4*>(a.in) + p * NF_PIXELS + k);
                if (!INV) { v.x *= a.temp; v.y *= a.temp; v.z *= a.temp; v.w *= a.temp; }   // noise_flow_model.py:501
            } else {
                v = philox_normal4(a.seed, a.offset, a.patch_base + (unsigned long long)p, (unsigned int)k);
                v.x *= a.temp; v.y *= a.temp; v.z *= a.temp; v.w *= a.temp;
            }
            S.z[k] = v;
            if (COND) S.y[k] = __ldg(reinterpret_cast<const float4*>(a.y) + p * NF_PIXELS + k);
        }
        __syncthreads();
        float ldj = 0.f;
        const int n_l = a.last_layer - a.first_layer;
        for (int step = 0; step < n_l; ++step) {
            const int l = INV ? a.first_layer + step : a.last_layer - 1 - step;
            const int op = prog.op[l];
            const float* pb = blob + prog.off[l];
            const bool probe = a.bn_stage != 0 && op == NF_KOP_COUPLING && l == (INV ? a.last_layer - 1 : a.first_layer);
            if (op == NF_KOP_COUPLING) {
                if (probe) {
                    if (a.bn_stage == 1) wide_coupling_any<W, 1, COND>(pb, S, warp, lane, INV, ldj);
                    else wide_coupling_any<W, 2, COND>(pb, S, warp, lane, INV, ldj);
                } else {
                    wide_coupling_any<W, 0, COND>(pb, S, warp, lane, INV, ldj);
                }
            } else if (op == NF_KOP_MIX) {
                for (int k = threadIdx.x; k < NF_PIXELS; k += WIDE_THREADS) S.z[k] = w_mix(S.z[k], pb + (INV ? 0 : 16));
            } else if (op == NF_KOP_SDN) {
                const float sa = pb[row * 4], sb = pb[row * 4 + 1];
                const float4* yp = reinterpret_cast<const float4*>(a.y) + p * NF_PIXELS;
                float acc = 0.f;
                for (int k = threadIdx.x; k < NF_PIXELS; k += WIDE_THREADS) {
                    const float4 y = __ldg(yp + k);
                    float4 z = S.z[k];
                    const float v0 = fmaf(sa, y.x, sb), v1 = fmaf(sa, y.y, sb), v2 = fmaf(sa, y.z, sb), v3 = fmaf(sa, y.w, sb);
                    const float r0 = rsqrtf(v0), r1 = rsqrtf(v1), r2 = rsqrtf(v2), r3 = rsqrtf(v3);
                    if (INV) { z.x *= r0; z.y *= r1; z.z *= r2; z.w *= r3; }                       // SdnEx5.py:125-126
                    else     { z.x *= v0 * r0; z.y *= v1 * r1; z.z *= v2 * r2; z.w *= v3 * r3; }   // SdnEx5.py:106-107
                    acc += (__logf(v0) + __logf(v1)) + (__logf(v2) + __logf(v3));
                    S.z[k] = z;
                }
                ldj += INV ? -0.5f * acc : 0.5f * acc;
            } else if (op == NF_KOP_GAIN) {
                const float mlt = INV ? pb[row * 4 + 1] : pb[row * 4];
                for (int k = threadIdx.x; k < NF_PIXELS; k += WIDE_THREADS) {
                    float4 z = S.z[k];
                    z.x *= mlt; z.y *= mlt; z.z *= mlt; z.w *= mlt;
                    S.z[k] = z;
                }
                if (threadIdx.x == 0) ldj += INV ? pb[row * 4 + 2] : -pb[row * 4 + 2];
            }
            __syncthreads();
        }
        if (a.bn_stage != 0) {   // probe launch: publish this patch's per-channel sums, nothing else
            __syncthreads();
            for (int k = threadIdx.x; k < 2 * W; k += WIDE_THREADS) {
                atomicAdd(a.bn_stats + k, (double)S.sacc[k]);
                S.sacc[k] = 0.f;
            }
            __syncthreads();
            continue;
        }
        // ---- epilogue: store the patch, reduce log-det / prior / latent statistics
        float s1 = 0.f, s2 = 0.f;
        float4* dst = a.out ? reinterpret_cast<float4*>(a.out) + p * NF_PIXELS : nullptr;
        for (int k = threadIdx.x; k < NF_PIXELS; k += WIDE_THREADS) {
            const float4 z = S.z[k];
            if (dst) __stcs(dst + k, z);
            s1 += (z.x + z.y) + (z.z + z.w);
            s2 = fmaf(z.x, z.x, fmaf(z.y, z.y, fmaf(z.z, z.z, fmaf(z.w, z.w, s2))));
        }
        const float t_ldj = w_cta_sum(ldj, S.red, warp, lane);
        const float t1 = w_cta_sum(s1, S.red, warp, lane);
        const float t2 = w_cta_sum(s2, S.red, warp, lane);
        if (threadIdx.x == 0) {
            const float logdet = t_ldj + (INV ? a.ldj_const : -a.ldj_const) + (a.logdet_in ? a.logdet_in[p] : 0.f);
            if (a.logdet) a.logdet[p] = logdet;
            if (a.nll) {   // -(logdet + sum -0.5 (log 2pi + z^2))       noise_flow_model.py:474-475,537-539
                const float logp = -0.5f * (NF_DIMS * 1.8378770664093453f + t2);
                a.nll[p] = -(logdet + logp);
            }
            if (a.sdz) {   // population std-dev of z                     noise_flow_model.py:477-478
                const float mean = t1 * (1.f / NF_DIMS);
                a.sdz[p] = sqrtf(fmaxf(t2 * (1.f / NF_DIMS) - mean * mean, 0.f));
            }
        }
        __syncthreads();
    }
}


}  // namespace nf
