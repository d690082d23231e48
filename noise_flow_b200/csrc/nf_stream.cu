// HBM-bound streaming kernel for chains made only of scale layers (sdn* / gain*), e.g. the reference's
// "sdn5|gain4" baseline arch (job_noise_flow.sh:53).  No spatial coupling -> no shared memory: one warp
// streams one patch with coalesced 512-byte row loads of x and y (16 LDG.128 in flight per lane), applies
// every scale layer in registers and reduces log-det / prior / latent statistics with warp shuffles.
// Roofline: 32 KiB read (+16 KiB if z is written) per patch and ~60 flop -> pure HBM bound.
//
// Reference semantics: AffineCouplingSdnEx5.py:66-132, AffineCouplingGainEx4.py:62-127,
// noise_flow_model.py:458-480 (loss), :525-541 (prior).
#include <cuda_runtime.h>
#include "nf_kernels.h"
#include "nf_params.h"

namespace nf {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

template <bool INV>
__global__ void __launch_bounds__(256)
nf_scale_stream_kernel(const __grid_constant__ NfModelParams mp, const NfChainArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < a.n; p += nwarps) {
        int row = a.rows ? a.rows[p] : a.default_row;
        row = min(max(row, 0), NF_MAX_ROWS - 1);
        // gather this patch's layer scalars: up to NF_MAX_SCALE sdn (a,b) pairs, one combined gain factor
        float sa[NF_MAX_SCALE], sb[NF_MAX_SCALE];
        int n_sdn = 0;
        float gmul = 1.f, ldj_gain = 0.f;
        for (int l = a.first_layer; l < a.last_layer; ++l) {
            const int slot = mp.slot[l];
            if (mp.op[l] == NF_KOP_SDN) {
#pragma unroll
                for (int k = 0; k < NF_MAX_SCALE; ++k)
                    if (k == n_sdn) { sa[k] = mp.sc[slot].t[row][0]; sb[k] = mp.sc[slot].t[row][1]; }
                ++n_sdn;
            } else if (mp.op[l] == NF_KOP_GAIN) {
                gmul *= INV ? mp.sc[slot].t[row][1] : mp.sc[slot].t[row][0];
                ldj_gain += mp.sc[slot].t[row][2];
            }
        }
        const float4* xin = a.in ? reinterpret_cast<const float4*>(a.in) + p * NF_PIXELS : nullptr;
        const float4* yin = (a.y && n_sdn) ? reinterpret_cast<const float4*>(a.y) + p * NF_PIXELS : nullptr;
        float4* dst = a.out ? reinterpret_cast<float4*>(a.out) + p * NF_PIXELS : nullptr;
        float lsum = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {
            float4 xv[8], yv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int idx = (r0 + j) * 32 + lane;
                xv[j] = __ldcs(xin + idx);
                yv[j] = yin ? __ldcs(yin + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 z = xv[j];
                if (!INV) { z.x *= a.temp; z.y *= a.temp; z.z *= a.temp; z.w *= a.temp; }
                float m0 = gmul, m1 = gmul, m2 = gmul, m3 = gmul;
#pragma unroll
                for (int k = 0; k < NF_MAX_SCALE; ++k) {
                    if (k < n_sdn) {
                        const float v0 = fmaf(sa[k], yv[j].x, sb[k]), v1 = fmaf(sa[k], yv[j].y, sb[k]);
                        const float v2 = fmaf(sa[k], yv[j].z, sb[k]), v3 = fmaf(sa[k], yv[j].w, sb[k]);
                        const float q0 = rsqrtf(v0), q1 = rsqrtf(v1), q2 = rsqrtf(v2), q3 = rsqrtf(v3);
                        if (INV) { m0 *= q0; m1 *= q1; m2 *= q2; m3 *= q3; }
                        else     { m0 *= v0 * q0; m1 *= v1 * q1; m2 *= v2 * q2; m3 *= v3 * q3; }
                        lsum += (__logf(v0) + __logf(v1)) + (__logf(v2) + __logf(v3));
                    }
                }
                z.x *= m0; z.y *= m1; z.z *= m2; z.w *= m3;
                if (dst) __stcs(dst + (r0 + j) * 32 + lane, z);
                s1 += (z.x + z.y) + (z.z + z.w);
                s2 = fmaf(z.x, z.x, fmaf(z.y, z.y, fmaf(z.z, z.z, fmaf(z.w, z.w, s2))));
            }
        }
        float ldj = wsum(INV ? -0.5f * lsum : 0.5f * lsum) + (INV ? ldj_gain : -ldj_gain);
        if (a.nll || a.sdz) { s1 = wsum(s1); s2 = wsum(s2); }
        if (lane == 0) {
            const float logdet = ldj + (INV ? a.ldj_const : -a.ldj_const);
            if (a.logdet) a.logdet[p] = logdet;
            if (a.nll) a.nll[p] = -(logdet - 0.5f * (NF_DIMS * 1.8378770664093453f + s2));
            if (a.sdz) {
                const float mean = s1 * (1.f / NF_DIMS);
                a.sdz[p] = sqrtf(fmaxf(s2 * (1.f / NF_DIMS) - mean * mean, 0.f));
            }
        }
    }
}

bool program_is_scale_only(const NfModelParams& mp, int first, int last) {
    int n_sdn = 0;
    for (int l = first; l < last; ++l) {
        if (mp.op[l] == NF_KOP_SDN) ++n_sdn;
        else if (mp.op[l] != NF_KOP_GAIN) return false;
    }
    return n_sdn <= NF_MAX_SCALE;
}

cudaError_t launch_scale_stream(const NfModelParams& mp, const NfChainArgs& args, bool inverse, int num_sms,
                                cudaStream_t stream) {
    if (args.n <= 0) return cudaSuccess;
    if (!args.in) return cudaErrorInvalidValue;   // in-kernel RNG lives in the chain kernel only
    long long ctas = (args.n + 7) / 8;
    const long long cap = (long long)num_sms * 8;
    if (ctas > cap) ctas = cap;
    if (inverse) nf_scale_stream_kernel<true><<<(unsigned)ctas, 256, 0, stream>>>(mp, args);
    else         nf_scale_stream_kernel<false><<<(unsigned)ctas, 256, 0, stream>>>(mp, args);
    return cudaGetLastError();
}

}  // namespace nf

// ================================================================================================
// Evaluation metrics that sit right after the hot path in the reference's drivers:
//   * Gaussian / camera-NLF NLL baselines  (sidd/PatchStatsCalculator.py:92-123, calc_baselines)
//   * histograms for the marginal KL divergence (sidd/sidd_utils.py:1044-1052, 1266-1274)
// Both are single streaming passes: HBM-bound, no shared state between patches except the histogram.
// ================================================================================================
namespace nf {

// nll_gauss[p] = sum 0.5*(log 2pi + log vg + x^2/vg),  nll_sdn[p] = same with v = y*nlf0 + nlf1
__global__ void __launch_bounds__(256)
nf_baseline_nll_kernel(const float4* __restrict__ x, const float4* __restrict__ y, float nlf0, float nlf1, float var_gauss,
                       long long n, float* __restrict__ nll_gauss, float* __restrict__ nll_sdn) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float log_vg = __logf(var_gauss), inv_vg = 1.f / var_gauss;
    for (long long p = warp; p < n; p += nwarps) {
        float sx2 = 0.f, ssdn = 0.f;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {
            float4 xv[8], yv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                xv[j] = __ldcs(x + p * NF_PIXELS + (r0 + j) * 32 + lane);
                yv[j] = __ldcs(y + p * NF_PIXELS + (r0 + j) * 32 + lane);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xs[4] = {xv[j].x, xv[j].y, xv[j].z, xv[j].w}, ys[4] = {yv[j].x, yv[j].y, yv[j].z, yv[j].w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float v = fmaf(ys[c], nlf0, nlf1), x2 = xs[c] * xs[c];
                    sx2 += x2;
                    ssdn += __logf(v) + __fdividef(x2, v);
                }
            }
        }
        sx2 = wsum(sx2);
        ssdn = wsum(ssdn);
        if (lane == 0) {
            const float c = NF_DIMS * 1.8378770664093453f;
            if (nll_gauss) nll_gauss[p] = 0.5f * (c + NF_DIMS * log_vg + sx2 * inv_vg);
            if (nll_sdn) nll_sdn[p] = 0.5f * (c + ssdn);
        }
    }
}

// counts[b] += #{ data in [edges[b], edges[b+1]) }, last bin closed on the right (np.histogram semantics);
// comparisons in double against the caller's double edges -> bit-identical bin decisions to numpy.
__global__ void __launch_bounds__(256)
nf_histogram_kernel(const float* __restrict__ data, long long count, const double* __restrict__ edges, int n_bins,
                    unsigned long long* __restrict__ counts) {
    extern __shared__ unsigned int sh_hist[];
    double* sh_edges = reinterpret_cast<double*>(sh_hist + ((n_bins + 1) & ~1));
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) sh_hist[i] = 0u;
    for (int i = threadIdx.x; i <= n_bins; i += blockDim.x) sh_edges[i] = edges[i];
    __syncthreads();
    const double lo = sh_edges[0], hi = sh_edges[n_bins];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const double v = (double)data[i];
        if (!(v >= lo) || !(v <= hi)) continue;      // outside the range (or NaN): not counted
        int a = 0, b = n_bins;                        // largest a with edges[a] <= v
        while (b - a > 1) {
            const int m = (a + b) >> 1;
            if (v >= sh_edges[m]) a = m; else b = m;
        }
        atomicAdd(&sh_hist[a], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
        if (sh_hist[i]) atomicAdd(&counts[i], (unsigned long long)sh_hist[i]);
}

cudaError_t launch_baseline_nll(const float* x, const float* y, float nlf0, float nlf1, float var_gauss, long long n,
                                float* nll_gauss, float* nll_sdn, int num_sms, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    long long ctas = (n + 7) / 8;
    if (ctas > (long long)num_sms * 8) ctas = (long long)num_sms * 8;
    nf_baseline_nll_kernel<<<(unsigned)ctas, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(y),
                                                               nlf0, nlf1, var_gauss, n, nll_gauss, nll_sdn);
    return cudaGetLastError();
}

cudaError_t launch_histogram(const float* data, long long count, const double* edges, int n_bins, unsigned long long* counts,
                             int num_sms, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    long long ctas = (count + 256 * 16 - 1) / (256 * 16);
    if (ctas > (long long)num_sms * 8) ctas = (long long)num_sms * 8;
    const size_t smem = (size_t)((n_bins + 1) & ~1) * 4 + (size_t)(n_bins + 1) * 8;
    nf_histogram_kernel<<<(unsigned)ctas, 256, smem, stream>>>(data, count, edges, n_bins, counts);
    return cudaGetLastError();
}

}  // namespace nf
