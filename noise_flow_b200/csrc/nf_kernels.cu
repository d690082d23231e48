// Fused Noise Flow bijector-chain kernels for sm_100a (B200).
//
// Design:  ONE WARP OWNS ONE 32x32x4 PATCH.  The patch (16 KiB, NHWC float4 per
// pixel) stays resident in that warp's shared memory across the whole bijector chain; lane l owns
// image column l and the warp streams down the 32 rows.  A coupling layer
// (1x1 mix -> conv3x3 -> BN -> ReLU -> conv1x1 -> BN -> ReLU -> edge-padded conv3x3 -> tanh/exp affine)
// is ONE software-pipelined pass over the rows: stage A (row t) mixes channels and publishes the
// conditioning half x0, stage B (row t-1) scatters that row into the three pending conv-1
// accumulators and emits h2 row t-2, stage C (row t-3) scatters the h2 row into the three pending
// conv-3 accumulators and finishes output row t-4 (affine update + log-det).  Rows travel between
// lanes through two tiny double-buffered row rings, so a pass needs one __syncwarp per row, no
// __syncthreads at all, and every activation is read from shared memory exactly once per consumer
// lane.  All convolution arithmetic is packed fma.rn.f32x2 (FFMA2) over input-channel pairs with
// the weights broadcast from the __grid_constant__ parameter block through uniform registers.
//
// Reference semantics restated here (file:line in /root/reference):
//   borealisflows/layers.py:117-130   Conv2d1x1 inverse/forward + constant log-det
//   borealisflows/layers.py:333-375   AffineCoupling forward/inverse + log-det
//   borealisflows/layers.py:452-498   real_nvp_conv_template (BN folded for is_training=False)
//   borealisflows/layers.py:555-583,651-674  add_edge_padding + conv2d_zeros
//   borealisflows/noise_flow_layers/AffineCouplingSdnEx5.py:66-132, AffineCouplingGainEx4.py:62-127
//   borealisflows/noise_flow_model.py:394-447,458-480,525-541  inverse / forward / prior / sd_z
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_params.h"
#include "nf_kernels.h"
#include "nf_rng.cuh"
#include "nf_chain_dev.cuh"

namespace nf {

// Per-warp shared memory (ZStore in nf_chain_dev.cuh says where the resident patch itself lives: with NF_Z_IN_TMEM = 1
// the 256 KB of tensor memory hold 16 patches per SM, shared memory only carries the two small row rings, so 16 warps
// are resident per SM instead of the 12 that fit when z lives in shared memory (216 KB)).
struct __align__(16) WarpSmem {
#if !NF_Z_IN_TMEM
    float4 z[NF_PIXELS];   // z[row * 32 + lane]
#endif
    float4 hr[2][34];      // h2 row ring: [1..32] = columns, [0] and [33] = zero halo
    float2 xr[2][34];      // x0 row ring
};

static_assert(sizeof(WarpSmem) == NF_WARP_SMEM_BYTES, "WarpSmem size");

}  // namespace nf
#include "nf_coupling.cuh"
namespace nf {

// Couplings 0..NF_FAST_SLOTS-1 get a templated copy of the pass: every weight then has a
// compile-time offset in the parameter block (LDCU.128 with an immediate address feeding FFMA2
// uniform-register operands; a run-time slot index makes ptxas fall back to per-thread LDC).  One loop
// body per coupling means the resident warps must not drift apart: run_layer() ends with a
// __syncthreads so all warps of the CTA (= of the SM) execute the same body at the same time -- without
// it the 12 warps sat in 8 different bodies and the kernel was instruction-cache bound (ncu, round 1:
// stall_no_instruction 1.6 per issue, icc hit rate 83 %).
#define NF_FAST_SLOTS 8
template <bool INV>
__device__ __forceinline__ void coupling_dispatch(const NfModelParams& mp, WarpSmem& s, const ZStore& zs, int lane, float& ldj, int slot) {
    switch (slot) {
#define NF_CASE(K) case K: coupling_pass<INV>(mp.cp[K], s, zs, lane, ldj); break;
        NF_CASE(0) NF_CASE(1) NF_CASE(2) NF_CASE(3) NF_CASE(4) NF_CASE(5) NF_CASE(6) NF_CASE(7)
#undef NF_CASE
        default: coupling_pass<INV>(mp.cp[slot], s, zs, lane, ldj); break;   // run-time slot: per-thread LDC weight fetches
    }
}

template <bool INV>
__device__ __forceinline__ void run_layer(const NfModelParams& mp, const NfChainArgs& a, WarpSmem& s, const ZStore& zs,
                                          int lane, int l, long long p, int row, float& ldj, float* stats) {
    const int op = mp.op[l], slot = mp.slot[l];
    const bool probe = a.bn_stage != 0 && op == NF_KOP_COUPLING && l == (INV ? a.last_layer - 1 : a.first_layer);
    if (probe) {   // batch-statistics BatchNorm: measure, do not transform (single compact copy, run-time slot)
        if (a.bn_stage == 1) coupling_stats_pass<INV, 1>(mp.cp[slot], s, zs, lane, stats);
        else                 coupling_stats_pass<INV, 2>(mp.cp[slot], s, zs, lane, stats);
        __syncthreads();
        return;
    }
    switch (op) {
        case NF_KOP_COUPLING: coupling_dispatch<INV>(mp, s, zs, lane, ldj, slot); break;
        case NF_KOP_MIX: mix_pass<INV>(mp.mix[slot], zs); break;
        case NF_KOP_SDN:
            sdn_pass<INV>(reinterpret_cast<const float4*>(a.y) + p * NF_PIXELS, mp.sc[slot].t[row][0], mp.sc[slot].t[row][1], zs, lane, ldj);
            break;
        case NF_KOP_GAIN:
            gain_pass<INV>(mp.sc[slot].t[row][0], mp.sc[slot].t[row][1], mp.sc[slot].t[row][2], zs, lane, ldj);
            break;
        default: break;
    }
    __syncthreads();   // layer boundary: keeps the CTA's warps in the same loop body (I-cache), orders smem
}

// ------------------------------------------------------------------------------------------------
// the fused chain kernel: persistent warps, one patch per warp at a time
// ------------------------------------------------------------------------------------------------
template <bool INV>
__global__ void __launch_bounds__(NF_MAX_CTA_THREADS, 1)
nf_chain_kernel(const __grid_constant__ NfModelParams mp, const NfChainArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    WarpSmem& s = reinterpret_cast<WarpSmem*>(smem_raw)[warp];

#if NF_Z_IN_TMEM
    __shared__ uint32_t tmem_base_smem;
    if (warp == 0) {   // the whole tensor memory of this SM: 512 columns x 128 lanes = 16 resident patches
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const ZStore zs = {tmem_base_smem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128)};
#else
    const ZStore zs = {s.z, lane};
#endif
    if (lane < 2) {   // zero halo columns of the row rings (never written again)
        s.hr[0][lane * 33] = s.hr[1][lane * 33] = make_float4(0.f, 0.f, 0.f, 0.f);
        s.xr[0][lane * 33] = s.xr[1][lane * 33] = make_float2(0.f, 0.f);
    }
    __syncwarp();

    // The patch loop runs on a CTA-uniform counter so that ptxas keeps the layer program and the weight
    // fetches on the uniform datapath; trailing warps without a patch of their own recompute the last
    // patch and only their stores are predicated off.
    const long long stride = (long long)gridDim.x * warps_per_cta;
    for (long long base = (long long)blockIdx.x * warps_per_cta; base < a.n; base += stride) {
        const bool active = base + warp < a.n;
        const long long p = active ? base + warp : a.n - 1;
        int row = a.rows ? a.rows[p] : a.default_row;
        row = min(max(row, 0), NF_MAX_ROWS - 1);

        // ---- load the patch into this warp's shared memory (coalesced 512 B per row)
        if (a.in) {
            const float4* src = reinterpret_cast<const float4*>(a.in) + p * NF_PIXELS;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                float4 v = __ldcs(src + r * 32 + lane);
                if (!INV) { v.x *= a.temp; v.y *= a.temp; v.z *= a.temp; v.w *= a.temp; }   // noise_flow_model.py:501
                zs.store(r, v);
            }
        } else {
#pragma unroll 2
            for (int r = 0; r < 32; ++r) {
                float4 v = philox_normal4(a.seed, a.offset, a.patch_base + (unsigned long long)p, (unsigned int)(r * 32 + lane));
                v.x *= a.temp; v.y *= a.temp; v.z *= a.temp; v.w *= a.temp;
                zs.store(r, v);
            }
        }
        zs.commit();
        __syncwarp();

        float ldj = 0.f;
        float stats[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        // (every pass below starts from committed z: load phase, scale / mix passes commit at their end, and a
        //  coupling pass commits at the top of each step and is followed by the commit of the next pass or epilogue)
        if (INV) {
            for (int l = a.first_layer; l < a.last_layer; ++l) run_layer<true>(mp, a, s, zs, lane, l, p, row, ldj, stats);
        } else {
            for (int l = a.last_layer - 1; l >= a.first_layer; --l) run_layer<false>(mp, a, s, zs, lane, l, p, row, ldj, stats);
        }

        if (a.bn_stage != 0) {   // probe launch: publish this patch's per-channel sums, nothing else
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float v = warp_sum(stats[k]);
                if (lane == 0 && active) atomicAdd(a.bn_stats + k, (double)v);
            }
            __syncwarp();
            continue;
        }
        // ---- epilogue: store the patch, reduce log-det / prior / latent statistics
        zs.commit();
        float s1 = 0.f, s2 = 0.f;
        float4* dst = (a.out && active) ? reinterpret_cast<float4*>(a.out) + p * NF_PIXELS : nullptr;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const float4 z = zs.load(r);
            if (dst) __stcs(dst + r * 32 + lane, z);
            s1 += (z.x + z.y) + (z.z + z.w);
            s2 = fmaf(z.x, z.x, fmaf(z.y, z.y, fmaf(z.z, z.z, fmaf(z.w, z.w, s2))));
        }
        ldj = warp_sum(ldj);
        if (a.nll || a.sdz) { s1 = warp_sum(s1); s2 = warp_sum(s2); }
        if (lane == 0 && active) {
            const float logdet = ldj + (INV ? a.ldj_const : -a.ldj_const) + (a.logdet_in ? a.logdet_in[p] : 0.f);
            if (a.logdet) a.logdet[p] = logdet;
            if (a.nll) {   // -(logdet + sum -0.5 (log 2pi + z^2))       noise_flow_model.py:474-475,537-539
                const float logp = -0.5f * (NF_DIMS * 1.8378770664093453f + s2);
                a.nll[p] = -(logdet + logp);
            }
            if (a.sdz) {   // population std-dev of z                     noise_flow_model.py:477-478
                const float mean = s1 * (1.f / NF_DIMS);
                a.sdz[p] = sqrtf(fmaxf(s2 * (1.f / NF_DIMS) - mean * mean, 0.f));
            }
        }
        __syncwarp();
    }
#if NF_Z_IN_TMEM
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_smem), "r"(512u) : "memory");
#endif
}

// ------------------------------------------------------------------------------------------------
// deterministic batch reduction: sums[0] = sum nll, sums[1] = sum sd_z, sums[2] = n   (fp64)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
nf_reduce_kernel(const float* __restrict__ nll, const float* __restrict__ sdz, long long n, double* __restrict__ sums) {
    __shared__ double sh[2][1024];
    double a = 0.0, b = 0.0;
    for (long long i = threadIdx.x; i < n; i += 1024) {
        if (nll) a += (double)nll[i];
        if (sdz) b += (double)sdz[i];
    }
    sh[0][threadIdx.x] = a;
    sh[1][threadIdx.x] = b;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + w];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { sums[0] = sh[0][0]; sums[1] = sh[1][0]; sums[2] = (double)n; }
}

// ------------------------------------------------------------------------------------------------
// squeeze2d / unsqueeze2d (reference borealisflows/utils.py:30-86): pure index permutation, bit-exact
// ------------------------------------------------------------------------------------------------
// out[n, h/f, w/f, chan] with chan = c*f*f + dy*f + dx  <- in[n, hh, ww, c]
//   chessboard: hh = (h/f)*f + dy, ww = (w/f)*f + dx       (utils.py:44-47)
//   patch     : hh = dy*(H/f) + h/f, ww = dx*(W/f) + w/f   (utils.py:48-51)
__global__ void nf_squeeze_kernel(const float* __restrict__ in, float* __restrict__ out, long long total,
                                  int H, int W, int C, int f, int patch_type, int inverse) {
    const int Ho = H / f, Wo = W / f, Co = C * f * f;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long t = idx;
        const int chan = (int)(t % Co); t /= Co;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho); t /= Ho;
        const long long n = t;
        const int dx = chan % f, dy = (chan / f) % f, c = chan / (f * f);
        const int hh = patch_type ? dy * Ho + ho : ho * f + dy;
        const int ww = patch_type ? dx * Wo + wo : wo * f + dx;
        const long long full = ((n * H + hh) * W + ww) * C + c;
        if (inverse) out[full] = in[idx]; else out[idx] = in[full];
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers (called by nf_api.cu)
// ------------------------------------------------------------------------------------------------
cudaError_t launch_chain(const NfModelParams& mp, const NfChainArgs& args, bool inverse, int num_sms, int warps_per_cta,
                         cudaStream_t stream) {
    if (args.n <= 0) return cudaSuccess;
    // CTA shape against wave quantisation: every SM takes part and the CTA is sized so that its rounds are full --
    // ceil(n / SMs) patches per SM in ceil(per_sm / warps_per_cta) rounds of equal size.  4 096 patches: two rounds of 14
    // warps on all 148 SMs (16-warp CTAs: a full round + a round with 40 SMs idle); 1 024 patches: one round of 7 warps on
    // every SM (16-warp CTAs: 64 SMs).  The kernel itself is untouched: blockDim decides how many patches a CTA keeps resident.
    {
        const long long g = args.n < (long long)num_sms ? args.n : (long long)num_sms;
        const long long per_sm = (args.n + g - 1) / g, rounds = (per_sm + warps_per_cta - 1) / warps_per_cta;
        warps_per_cta = (int)((per_sm + rounds - 1) / rounds);
    }
    const size_t smem = (size_t)warps_per_cta * sizeof(WarpSmem);
    static bool attr_done[NF_MAX_DEVICES][2] = {};   // per device; idempotent: a benign race sets the same value twice
    const int dev = device_slot();
    cudaError_t e;
    if (inverse) {
        if (!attr_done[dev][0]) {
            e = cudaFuncSetAttribute(nf_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NF_MAX_CTA_SMEM);
            if (e != cudaSuccess) return e;
            attr_done[dev][0] = true;
        }
    } else if (!attr_done[dev][1]) {
        e = cudaFuncSetAttribute(nf_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NF_MAX_CTA_SMEM);
        if (e != cudaSuccess) return e;
        attr_done[dev][1] = true;
    }
    long long ctas = (args.n + warps_per_cta - 1) / warps_per_cta;
    if (ctas > num_sms) ctas = num_sms;
    if (inverse) nf_chain_kernel<true><<<(unsigned)ctas, warps_per_cta * 32, smem, stream>>>(mp, args);
    else         nf_chain_kernel<false><<<(unsigned)ctas, warps_per_cta * 32, smem, stream>>>(mp, args);
    return cudaGetLastError();
}

cudaError_t launch_reduce(const float* nll, const float* sdz, long long n, double* sums, cudaStream_t stream) {
    nf_reduce_kernel<<<1, 1024, 0, stream>>>(nll, sdz, n, sums);
    return cudaGetLastError();
}

cudaError_t launch_squeeze(const float* in, float* out, long long n, int H, int W, int C, int factor, int patch_type,
                           int inverse, cudaStream_t stream) {
    const long long total = n * H * W * C;
    if (total <= 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    nf_squeeze_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, out, total, H, W, C, factor, patch_type, inverse);
    return cudaGetLastError();
}

}  // namespace nf
