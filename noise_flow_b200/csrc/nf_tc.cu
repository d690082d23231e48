// Tensor-core formulation of the Noise Flow chain for sm_100a: the two 3x3 convolutions of every coupling net
// run on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), everything else stays fp32 on the
// CUDA cores.  Selected with nf_model_set_launch(..)/NF_TC (see nf_api.cu); parity-tested like the fp32 kernel.
//
// Why this is a dense contraction after all.  A 3x3 conv with 4 output channels is a GEMM with N = 4 -- far off
// the tensor cores' design point -- but the UMMA shared-memory descriptor lets the A operand be read IN PLACE
// from a pixel-major image: one pixel = 16 bytes = 8 bf16 = one row of a K-major core matrix, 8 consecutive
// pixels = one core matrix (SBO = 128 B), and the second K chunk of an MMA may start at ANY 16-byte offset
// (LBO), i.e. at another tap of the stencil.  With the patch stored as a zero-padded 34-wide image, one
// M=128,N=16,K=16 MMA therefore computes two stencil taps for 128 consecutive pixels with no im2col at all
// (validated by tools/tc_probe.cu).  fp32 accuracy comes from bf16 hi/lo splitting: a pixel's 8 K-slots hold
// (hi, lo) pairs of its channels, the B tile holds W_hi in columns 0-3 (applied to hi and lo slots) and W_lo in
// columns 4-7 (applied to hi slots), the epilogue adds column j and j+4:  a*W - a_lo*W_lo, relative error 2^-16.
//
// Work decomposition: one GROUP of 4 warps (128 threads = the 128 TMEM lanes) owns one patch at a time; a CTA
// holds NF_TC_GROUPS groups (one CTA per SM) so that one group's CUDA-core epilogue overlaps another group's
// MMAs.  Per coupling and patch: prep (1x1 mix, publish x0 image) -> 45 MMAs (conv-1, 9 row tiles x 5 tap pairs)
// -> epilogue 1 (TMEM -> +bias via a constant-one K slot, ReLU, 1x1 conv, ReLU, publish h2 image) -> 45 MMAs
// (conv-3) -> epilogue 2 (edge bias, tanh/exp affine update of z, log-det).  B tiles for all couplings are built
// once per CTA, in shared memory, from the fp32 parameter block.
//
// Reference semantics: identical to nf_kernels.cu / nf_coupling.cuh (layers.py:117-130, 333-375, 452-498, 555-583,
// 651-674; noise_flow_model.py:394-480).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_kernels.h"
#include "nf_params.h"

namespace nf {

constexpr int TC_GROUPS = NF_TC_GROUPS;
constexpr int TC_IMG = 1224;             // padded-linear positions 0..1221 (34-wide rows + tap overreach), /8
constexpr int TC_P0 = 35;                // position of pixel (0,0): row 1, column 1 of the padded map
constexpr int TC_TILES = 9;              // 9 x 128 positions cover rows 1..32 of the padded map
constexpr int TC_COLS = 16 * TC_TILES;   // TMEM columns per group (N = 16 per tile)
constexpr int TC_WB_PER_SLOT = 10 * 32;  // uint4 per coupling: 10 B tiles (5 conv-1, 5 conv-3) x 512 B

struct __align__(16) TcGroupSmem {
    float4 z[NF_PIXELS];
    uint4 img[TC_IMG];
    float red[16];
    uint64_t mbar;
    uint64_t pad_;
};

struct __align__(16) TcSmem {
    uint4 wb[NF_TC_SLOTS * TC_WB_PER_SLOT];
    TcGroupSmem grp[TC_GROUPS];
    uint32_t tmem_base;
    uint32_t pad_[3];
};
static_assert(sizeof(TcSmem) <= 227 * 1024, "TcSmem too large");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);   // version 1, SWIZZLE_NONE
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N = 16, M = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tNF_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra NF_DONE;\n\tbra NF_WAIT;\n\tNF_DONE:\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void group_barrier(int g) { asm volatile("bar.sync %0, 128;" :: "r"(g + 1) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&d)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = __uint_as_float(r[k]);
}

// (hi, lo) bf16 split of two floats: hi = rn(v), lo = rn(v - hi)
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float r0 = v0 - __uint_as_float(hi << 16), r1 = v1 - __uint_as_float(hi & 0xFFFF0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint16_t bf16_bits(float v) {
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    return *reinterpret_cast<const uint16_t*>(&b);
}

__device__ __forceinline__ float tc_tanh(float v) { return 1.f - __fdividef(2.f, exp2f(v * 2.885390081777927f) + 1.f); }
__device__ __forceinline__ float tc_exp(float v) { return exp2f(v * 1.4426950408889634f); }
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ float4 tc_mix(float4 v, const float (&m)[16]) {   // out[o] = sum_i v[i] * m[o*4+i]
    float r[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) r[o] = fmaf(v.x, m[o * 4], fmaf(v.y, m[o * 4 + 1], fmaf(v.z, m[o * 4 + 2], v.w * m[o * 4 + 3])));
    return make_float4(r[0], r[1], r[2], r[3]);
}

// ---- B tiles: element (n, kk) of tile j of coupling `P` -------------------------------------------------------
// conv-1 tiles j = 0..4, conv-3 tiles j = 5..9; tap pair (2j', 2j'+1), taps in row-major (dy, dx) order = ascending
// image offset; K slot layout of a pixel: conv-1 image [hi0, hi1, lo0, lo1, ONE, 0, 0, 0], conv-3 image
// [hi0..hi3, lo0..lo3].  Columns n 0-3: hi part of the weight (seen by hi and lo slots), 4-7: lo part (hi slots only).
__device__ __forceinline__ float tc_b_value(const NfCouplingP& P, int j, int n, int kk) {
    if (n >= 8) return 0.f;
    const bool conv3 = j >= 5;
    const int jj = conv3 ? j - 5 : j;
    const int tap = 2 * jj + (kk >> 3), slot = kk & 7, o = n & 3;
    const bool lo_col = n >= 4;
    if (tap > 8) return 0.f;
    float w;
    bool hi_slot;
    if (!conv3) {
        if (slot == 4) { if (tap != 4) return 0.f; w = P.b1[o]; hi_slot = true; }        // bias through the ONE slot
        else if (slot > 4) return 0.f;
        else { w = P.w1[tap / 3][tap % 3][o][slot & 1]; hi_slot = slot < 2; }
    } else {
        w = P.w3[tap / 3][tap % 3][o][slot & 3];
        hi_slot = slot < 4;
    }
    const float whi = __bfloat162float(__float2bfloat16_rn(w));
    if (!lo_col) return whi;                   // W_hi multiplies hi and lo slots
    return hi_slot ? w - whi : 0.f;            // W_lo multiplies hi slots only
}

// ---- one coupling on the tensor cores -------------------------------------------------------------------------
template <bool INV>
__device__ __forceinline__ void tc_coupling(const NfCouplingP& P, TcSmem& S, TcGroupSmem& G, const int g, const int i,
                                            const int slot, const uint32_t tmem_g, uint32_t& phase, float& ldj) {
    const bool has_mix = P.has_mix != 0;
    // ---------------- prep: (inverse) 1x1 mix in fp32, publish x0 = channels 0,1 as (hi, lo, ONE) pixels
    {
        float m[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) m[k] = (&P.a[0][0])[k];
#pragma unroll 4
        for (int k = 0; k < 8; ++k) {
            const int px = k * 128 + i, r = px >> 5, c = px & 31;
            float4 z = G.z[px];
            if (INV && has_mix) { z = tc_mix(z, m); G.z[px] = z; }
            uint32_t hi, lo;
            split2(z.x, z.y, hi, lo);
            G.img[(r + 1) * 34 + (c + 1)] = make_uint4(hi, lo, 0x00003F80u, 0u);
        }
    }
    fence_async_smem();
    group_barrier(g);
    const uint32_t img_addr = smem_u32(&G.img[0]);
    const uint32_t wb_addr = smem_u32(&S.wb[slot * TC_WB_PER_SLOT]);
    const uint32_t mbar = smem_u32(&G.mbar);
    // tap offsets (pixels) in ascending order; pair (2j, 2j+1); the 10th "tap" is a dummy with zero weights
    auto issue = [&](int conv) {
        tc_fence_after();
        // tap pairs outermost, row tiles innermost: consecutive MMAs hit different accumulators, so the
        // tensor pipe is not serialised on the accumulate dependency (46 cycles per dependent MMA, tools/tc_probe.cu)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int ta = 2 * j, tb = 2 * j + 1;
            const int offa = (ta / 3 - 1) * 34 + (ta % 3 - 1);
            const int offb = tb <= 8 ? (tb / 3 - 1) * 34 + (tb % 3 - 1) : offa;
            const uint64_t bd = make_desc(wb_addr + (uint32_t)(conv * 5 + j) * 512u, 128u, 256u);
#pragma unroll 1
            for (int t = 0; t < TC_TILES; ++t) {
                const uint64_t ad = make_desc(img_addr + (uint32_t)(TC_P0 + 128 * t + offa) * 16u, (uint32_t)(offb - offa) * 16u, 128u);
                mma_bf16(tmem_g + 16u * t, ad, bd, j > 0 ? 1u : 0u);
            }
        }
        mma_commit(mbar);
    };
    // ---------------- conv-1 on the tensor cores
    if (i == 0) issue(0);
    mbar_wait(mbar, phase);
    phase ^= 1u;
    tc_fence_after();
    // ---------------- epilogue 1: bias came through the ONE slot; ReLU; 1x1 conv + bias + ReLU (fp32); publish h2
    {
        float w2[16], b2[4];
#pragma unroll
        for (int k = 0; k < 16; ++k) w2[k] = (&P.w2[0][0])[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) b2[k] = P.b2[k];
        const uint32_t lane_addr = tmem_g + ((uint32_t)((i >> 5) * 32) << 16);
#pragma unroll 1
        for (int t = 0; t < TC_TILES; ++t) {
            float d[8];
            tmem_ld8(lane_addr + 16u * t, d);
            const int pos = TC_P0 + 128 * t + i;
            const int R = pos / 34, C = pos - R * 34;
            const bool valid = R >= 1 && R <= 32 && C >= 1 && C <= 32;
            float h1[4], h2[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) h1[o] = fmaxf(d[o] + d[o + 4], 0.f);
#pragma unroll
            for (int o = 0; o < 4; ++o)
                h2[o] = fmaxf(fmaf(h1[0], w2[o * 4], fmaf(h1[1], w2[o * 4 + 1], fmaf(h1[2], w2[o * 4 + 2], fmaf(h1[3], w2[o * 4 + 3], b2[o])))), 0.f);
            uint32_t hi01, lo01, hi23, lo23;
            split2(h2[0], h2[1], hi01, lo01);
            split2(h2[2], h2[3], hi23, lo23);
            if (pos < TC_IMG) G.img[pos] = valid ? make_uint4(hi01, hi23, lo01, lo23) : make_uint4(0u, 0u, 0u, 0u);
        }
    }
    tc_fence_before();
    fence_async_smem();
    group_barrier(g);
    // ---------------- conv-3 on the tensor cores
    if (i == 0) issue(1);
    mbar_wait(mbar, phase);
    phase ^= 1u;
    tc_fence_after();
    // ---------------- epilogue 2: edge-indicator bias, affine coupling update of z, log-det
    {
        float m[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) m[k] = (&P.ainv[0][0])[k];
        const float scale = P.scale;
        const uint32_t lane_addr = tmem_g + ((uint32_t)((i >> 5) * 32) << 16);
#pragma unroll 1
        for (int t = 0; t < TC_TILES; ++t) {
            float d[8];
            tmem_ld8(lane_addr + 16u * t, d);
            const int pos = TC_P0 + 128 * t + i;
            const int R = pos / 34, C = pos - R * 34;
            const bool valid = R >= 1 && R <= 32 && C >= 1 && C <= 32;
            if (valid) {
                const int rc = R == 1 ? 0 : (R == 32 ? 2 : 1), cc = C == 1 ? 0 : (C == 32 ? 2 : 1);
                float h3[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) h3[o] = d[o] + d[o + 4] + P.b3[rc][cc][o];
                const float ls0 = scale * tc_tanh(h3[2]), ls1 = scale * tc_tanh(h3[3]);
                const int px = (R - 1) * 32 + (C - 1);
                float4 z = G.z[px];
                if (INV) {
                    z.z = fmaf(z.z, tc_exp(ls0), h3[0]);
                    z.w = fmaf(z.w, tc_exp(ls1), h3[1]);
                    ldj += ls0 + ls1;
                } else {
                    z.z = (z.z - h3[0]) * tc_exp(-ls0);
                    z.w = (z.w - h3[1]) * tc_exp(-ls1);
                    ldj -= ls0 + ls1;
                    if (has_mix) z = tc_mix(z, m);
                }
                G.z[px] = z;
            }
        }
    }
    tc_fence_before();
    group_barrier(g);
}

template <bool INV>
__global__ void __launch_bounds__(TC_GROUPS * 128, 1)
nf_chain_tc_kernel(const __grid_constant__ NfModelParams mp, const NfChainArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw);
    const int tid = threadIdx.x, g = tid >> 7, i = tid & 127, lane = tid & 31;
    TcGroupSmem& G = S.grp[g];

    // ---- CTA prologue: B tiles for every coupling slot, zeroed images, barriers, tensor memory
    {
        int n_cp = 0;
        for (int l = 0; l < mp.n_layers; ++l) n_cp += mp.op[l] == NF_KOP_COUPLING;
        uint16_t* wb16 = reinterpret_cast<uint16_t*>(S.wb);
        for (int e = tid; e < n_cp * 10 * 256; e += blockDim.x) {
            const int s = e / 2560, rem = e - s * 2560, j = rem >> 8, n = (rem >> 4) & 15, kk = rem & 15;
            const float v = tc_b_value(mp.cp[s], j, n, kk);
            // canonical K-major, no swizzle: (n, kk) at (n%8)*16 + (n/8)*256 + (kk/8)*128 + (kk%8)*2 bytes
            wb16[(s * 2560) + j * 256 + ((n & 7) * 16 + (n >> 3) * 256 + (kk >> 3) * 128 + (kk & 7) * 2) / 2] = bf16_bits(v);
        }
        for (int k = i; k < TC_IMG; k += 128) G.img[k] = make_uint4(0u, 0u, 0u, 0u);
        if (i == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&G.mbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tmem_g = S.tmem_base + (uint32_t)(g * TC_COLS);
    uint32_t phase = 0;

    const long long stride = (long long)gridDim.x * TC_GROUPS;
    for (long long p = (long long)blockIdx.x * TC_GROUPS + g; p < a.n; p += stride) {
        int row = a.rows ? a.rows[p] : a.default_row;
        row = min(max(row, 0), NF_MAX_ROWS - 1);
        // ---- load the patch (8 pixels per thread, coalesced 512 B per warp)
        if (a.in) {
            const float4* src = reinterpret_cast<const float4*>(a.in) + p * NF_PIXELS;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float4 v = __ldcs(src + k * 128 + i);
                if (!INV) { v.x *= a.temp; v.y *= a.temp; v.z *= a.temp; v.w *= a.temp; }
                G.z[k * 128 + i] = v;
            }
        } else {
            for (int k = 0; k < 8; ++k) G.z[k * 128 + i] = make_float4(0.f, 0.f, 0.f, 0.f);   // Philox: chain kernel only
        }
        group_barrier(g);
        float ldj = 0.f;
        const int l0 = INV ? a.first_layer : a.last_layer - 1, l1 = INV ? a.last_layer : a.first_layer - 1, dl = INV ? 1 : -1;
        for (int l = l0; l != l1; l += dl) {
            const int op = mp.op[l], slot = mp.slot[l];
            if (op == NF_KOP_COUPLING) {
                tc_coupling<INV>(mp.cp[slot], S, G, g, i, slot, tmem_g, phase, ldj);
            } else if (op == NF_KOP_MIX) {
                float m[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) m[k] = INV ? (&mp.mix[slot].a[0][0])[k] : (&mp.mix[slot].ainv[0][0])[k];
                for (int k = 0; k < 8; ++k) G.z[k * 128 + i] = tc_mix(G.z[k * 128 + i], m);
                group_barrier(g);
            } else if (op == NF_KOP_SDN) {
                const float sa = mp.sc[slot].t[row][0], sb = mp.sc[slot].t[row][1];
                const float4* yp = reinterpret_cast<const float4*>(a.y) + p * NF_PIXELS;
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 y = __ldg(yp + k * 128 + i);
                    float4 z = G.z[k * 128 + i];
                    const float v0 = fmaf(sa, y.x, sb), v1 = fmaf(sa, y.y, sb), v2 = fmaf(sa, y.z, sb), v3 = fmaf(sa, y.w, sb);
                    const float r0 = rsqrtf(v0), r1 = rsqrtf(v1), r2 = rsqrtf(v2), r3 = rsqrtf(v3);
                    if (INV) { z.x *= r0; z.y *= r1; z.z *= r2; z.w *= r3; }
                    else     { z.x *= v0 * r0; z.y *= v1 * r1; z.z *= v2 * r2; z.w *= v3 * r3; }
                    acc += (__logf(v0) + __logf(v1)) + (__logf(v2) + __logf(v3));
                    G.z[k * 128 + i] = z;
                }
                ldj += INV ? -0.5f * acc : 0.5f * acc;
                group_barrier(g);
            } else if (op == NF_KOP_GAIN) {
                const float mlt = INV ? mp.sc[slot].t[row][1] : mp.sc[slot].t[row][0];
                for (int k = 0; k < 8; ++k) {
                    float4 z = G.z[k * 128 + i];
                    z.x *= mlt; z.y *= mlt; z.z *= mlt; z.w *= mlt;
                    G.z[k * 128 + i] = z;
                }
                if (i == 0) ldj += INV ? mp.sc[slot].t[row][2] : -mp.sc[slot].t[row][2];
                group_barrier(g);
            }
        }
        // ---- epilogue: store, reduce
        float s1 = 0.f, s2 = 0.f;
        float4* dst = a.out ? reinterpret_cast<float4*>(a.out) + p * NF_PIXELS : nullptr;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 z = G.z[k * 128 + i];
            if (dst) __stcs(dst + k * 128 + i, z);
            s1 += (z.x + z.y) + (z.z + z.w);
            s2 = fmaf(z.x, z.x, fmaf(z.y, z.y, fmaf(z.z, z.z, fmaf(z.w, z.w, s2))));
        }
        ldj = wsum(ldj); s1 = wsum(s1); s2 = wsum(s2);
        if (lane == 0) { G.red[(i >> 5) * 4] = ldj; G.red[(i >> 5) * 4 + 1] = s1; G.red[(i >> 5) * 4 + 2] = s2; }
        group_barrier(g);
        if (i == 0) {
            ldj = (G.red[0] + G.red[4]) + (G.red[8] + G.red[12]);
            s1 = (G.red[1] + G.red[5]) + (G.red[9] + G.red[13]);
            s2 = (G.red[2] + G.red[6]) + (G.red[10] + G.red[14]);
            const float logdet = ldj + (INV ? a.ldj_const : -a.ldj_const) + (a.logdet_in ? a.logdet_in[p] : 0.f);
            if (a.logdet) a.logdet[p] = logdet;
            if (a.nll) a.nll[p] = -(logdet - 0.5f * (NF_DIMS * 1.8378770664093453f + s2));
            if (a.sdz) {
                const float mean = s1 * (1.f / NF_DIMS);
                a.sdz[p] = sqrtf(fmaxf(s2 * (1.f / NF_DIMS) - mean * mean, 0.f));
            }
        }
        group_barrier(g);
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(S.tmem_base), "r"(512u) : "memory");
}

bool tc_program_supported(const NfModelParams& mp, const NfChainArgs& a) {
    int n_cp = 0;
    for (int l = 0; l < mp.n_layers; ++l) n_cp += mp.op[l] == NF_KOP_COUPLING;
    return n_cp >= 1 && n_cp <= NF_TC_SLOTS && a.in != nullptr && a.bn_stage == 0;
}

cudaError_t launch_chain_tc(const NfModelParams& mp, const NfChainArgs& args, bool inverse, int num_sms, cudaStream_t stream) {
    if (args.n <= 0) return cudaSuccess;
    static bool attr_done_dev[NF_MAX_DEVICES][2] = {};   // per device
    bool* attr_done = attr_done_dev[device_slot()];
    cudaError_t e;
    const int k = inverse ? 0 : 1;
    if (!attr_done[k]) {
        e = inverse ? cudaFuncSetAttribute(nf_chain_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem))
                    : cudaFuncSetAttribute(nf_chain_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem));
        if (e != cudaSuccess) return e;
        attr_done[k] = true;
    }
    long long ctas = (args.n + TC_GROUPS - 1) / TC_GROUPS;
    if (ctas > num_sms) ctas = num_sms;
    if (inverse) nf_chain_tc_kernel<true><<<(unsigned)ctas, TC_GROUPS * 128, sizeof(TcSmem), stream>>>(mp, args);
    else         nf_chain_tc_kernel<false><<<(unsigned)ctas, TC_GROUPS * 128, sizeof(TcSmem), stream>>>(mp, args);
    return cudaGetLastError();
}

}  // namespace nf
