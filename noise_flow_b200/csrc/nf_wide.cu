// Fused Noise Flow chain for coupling nets wider than the shipped width 4 (real_nvp_conv_template(width),
// layers.py:452-498; `--width` in sidd/ArgParser.py:43, "for Noise Flow it is 32": job_noise_flow.sh:19).
//
// ONE CTA OWNS ONE PATCH.  At width W the hidden activation of a coupling net is 34 x 34 x W floats (148 KB at
// W = 32) -- it no longer fits a warp's share of the SM, so the whole CTA (16 warps; warp w owns image rows
// w, w+16; lane = column) works on one patch that stays resident in shared memory across the whole chain:
//   z   [32][32] float4                        the patch
//   h2  [34][W/4][34] float4                   padded hidden image, channel groups of 4 as planes so that the 32
//                                              lanes of a warp read 32 consecutive float4 (conflict-free)
//   w   one coupling's folded parameters       staged from the model's device blob at every layer
// A coupling is two CTA-wide phases: P1 conv3x3(2->W)+BN+ReLU -> conv1x1(W->W)+BN+ReLU per pixel in registers,
// h2 to shared memory; P2 edge-padded conv3x3(W->4) -> tanh/exp affine update + log-det.  BatchNorm (moving
// statistics), exp(3 logs) and the edge-indicator channel are folded on the host exactly as for width 4
// (nf_api.cu fold_wide).  The binding roof is the FP32 pipe: 86 W + 18 W + W^2 + 36 W MAC per pixel and coupling.
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
    float4 z[NF_PIXELS];
    float4 h2[34 * (W / 4) * 34];
    float w[NfWideLayout<W>::SIZE > 128 ? NfWideLayout<W>::SIZE : 128];
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
template <int W, bool INV, int STAGE>
__device__ __forceinline__ void wide_coupling(const float* __restrict__ gblob, WideSmem<W>& S, int warp, int lane, float& ldj) {
    using L = NfWideLayout<W>;
    constexpr int G = W / 4;
    for (int k = threadIdx.x * 4; k < L::SIZE; k += WIDE_THREADS * 4)
        *reinterpret_cast<float4*>(&S.w[k]) = *reinterpret_cast<const float4*>(&gblob[k]);
    __syncthreads();
    const float* w = S.w;
    const bool has_mix = w[L::META] != 0.f;
    if (INV && has_mix) {
        for (int r = warp; r < 32; r += WIDE_WARPS) S.z[r * 32 + lane] = w_mix(S.z[r * 32 + lane], w + L::A);
        __syncthreads();
    }
    // Each thread owns TWO pixels (rows `warp` and `warp + 16`, column `lane`) and walks them together: every weight
    // fetched from shared memory (broadcast LDS.128) feeds both, which halves the shared-memory instructions per FMA
    // (with one pixel the LDS pipe, 1 per clock and SM, saturates together with the FMA pipe).
    static_assert(WIDE_WARPS == 16, "two rows per warp");
    const int rows2[2] = {warp, warp + 16};
    // ---- P1: conv 3x3 SAME (2 -> W) + folded BN + ReLU ; conv 1x1 (W -> W) + folded BN + ReLU
    {
        float h1[2][W];
#pragma unroll
        for (int o = 0; o < W; ++o) h1[0][o] = h1[1][o] = w[L::B1 + o];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int cc = lane + dx - 1;
                float2 x0[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int rr = rows2[q] + dy - 1;
                    x0[q] = make_float2(0.f, 0.f);
                    if (rr >= 0 && rr <= 31 && cc >= 0 && cc <= 31) { const float4 v = S.z[rr * 32 + cc]; x0[q] = make_float2(v.x, v.y); }
                }
                const float* wt = w + L::W1 + (dy * 3 + dx) * W * 2;
#pragma unroll
                for (int o = 0; o < W; o += 2) {
                    const float4 wv = *reinterpret_cast<const float4*>(wt + o * 2);   // (o,i0) (o,i1) (o+1,i0) (o+1,i1)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        h1[q][o] = fmaf(x0[q].x, wv.x, fmaf(x0[q].y, wv.y, h1[q][o]));
                        h1[q][o + 1] = fmaf(x0[q].x, wv.z, fmaf(x0[q].y, wv.w, h1[q][o + 1]));
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
            for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = w[L::B2 + o + j];
#pragma unroll
            for (int i = 0; i < W; i += 4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 wv = *reinterpret_cast<const float4*>(w + L::W2 + (o + j) * W + i);
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
        float pre[2][4];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int rc = rows2[q] == 0 ? 0 : (rows2[q] == 31 ? 2 : 1);
#pragma unroll
            for (int o = 0; o < 4; ++o) pre[q][o] = w[L::B3 + (rc * 3 + cc) * 4 + o];
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float* wt = w + L::W3 + (dy * 3 + dx) * W * 4;
#pragma unroll 4
                for (int g = 0; g < G; ++g) {
                    const float4 w0 = *reinterpret_cast<const float4*>(wt + (g * 4 + 0) * 4);
                    const float4 w1 = *reinterpret_cast<const float4*>(wt + (g * 4 + 1) * 4);
                    const float4 w2 = *reinterpret_cast<const float4*>(wt + (g * 4 + 2) * 4);
                    const float4 w3 = *reinterpret_cast<const float4*>(wt + (g * 4 + 3) * 4);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 h = S.h2[((rows2[q] + dy) * G + g) * 34 + lane + dx];
                        pre[q][0] = fmaf(h.x, w0.x, fmaf(h.y, w1.x, fmaf(h.z, w2.x, fmaf(h.w, w3.x, pre[q][0]))));
                        pre[q][1] = fmaf(h.x, w0.y, fmaf(h.y, w1.y, fmaf(h.z, w2.y, fmaf(h.w, w3.y, pre[q][1]))));
                        pre[q][2] = fmaf(h.x, w0.z, fmaf(h.y, w1.z, fmaf(h.z, w2.z, fmaf(h.w, w3.z, pre[q][2]))));
                        pre[q][3] = fmaf(h.x, w0.w, fmaf(h.y, w1.w, fmaf(h.z, w2.w, fmaf(h.w, w3.w, pre[q][3]))));
                    }
                }
            }
        const float scale = w[L::META + 1];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float ls0 = scale * w_tanh(pre[q][2]), ls1 = scale * w_tanh(pre[q][3]);      // layers.py:362-365
            float4 z = S.z[rows2[q] * 32 + lane];
            if (INV) {                                                                          // layers.py:355-375
                z.z = fmaf(z.z, w_exp(ls0), pre[q][0]);
                z.w = fmaf(z.w, w_exp(ls1), pre[q][1]);
                ldj += ls0 + ls1;
            } else {                                                                            // layers.py:333-353
                z.z = (z.z - pre[q][0]) * w_exp(-ls0);
                z.w = (z.w - pre[q][1]) * w_exp(-ls1);
                ldj -= ls0 + ls1;
            }
            if (!INV && has_mix) z = w_mix(z, w + L::AINV);
            S.z[rows2[q] * 32 + lane] = z;
        }
    }
}

template <int W, bool INV>
__global__ void __launch_bounds__(WIDE_THREADS, 1)
nf_wide_chain_kernel(const NfWideProgram prog, const float* __restrict__ blob, const NfChainArgs a) {
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
                    if (a.bn_stage == 1) wide_coupling<W, INV, 1>(pb, S, warp, lane, ldj);
                    else wide_coupling<W, INV, 2>(pb, S, warp, lane, ldj);
                } else {
                    wide_coupling<W, INV, 0>(pb, S, warp, lane, ldj);
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

template <int W>
static cudaError_t launch_wide_w(const NfWideProgram& prog, const float* blob, const NfChainArgs& a, bool inverse, int num_sms,
                                 cudaStream_t stream) {
    const size_t smem = sizeof(WideSmem<W>);
    static bool attr_done_dev[NF_MAX_DEVICES] = {};   // per device (and per width: one instance of this template each)
    bool& attr_done = attr_done_dev[device_slot()];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(nf_wide_chain_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(nf_wide_chain_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const long long grid = a.n < (long long)num_sms ? a.n : (long long)num_sms;   // persistent: one CTA per SM
    if (inverse) nf_wide_chain_kernel<W, true><<<(unsigned)grid, WIDE_THREADS, smem, stream>>>(prog, blob, a);
    else nf_wide_chain_kernel<W, false><<<(unsigned)grid, WIDE_THREADS, smem, stream>>>(prog, blob, a);
    return cudaGetLastError();
}

bool wide_width_supported(int width) { return width == 8 || width == 16 || width == 32; }

cudaError_t launch_chain_wide(const NfWideProgram& prog, const float* blob, const NfChainArgs& a, bool inverse, int num_sms,
                              cudaStream_t stream) {
    switch (prog.width) {
        case 8: return launch_wide_w<8>(prog, blob, a, inverse, num_sms, stream);
        case 16: return launch_wide_w<16>(prog, blob, a, inverse, num_sms, stream);
        case 32: return launch_wide_w<32>(prog, blob, a, inverse, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace nf
