// All-fp32 fused Noise Flow chain with the two 3x3 convolutions of every coupling net in a VERTICAL WINOGRAD F(2,3)
// form ("Winograd" chain kernel; the default chain kernel since round 5).  Same organisation as nf_kernels.cu -- one warp owns one 32x32x4
// patch resident in tensor memory, lane = image column, packed fma.rn.f32x2 arithmetic, weights through the uniform
// datapath -- but a coupling pass walks the patch TWO rows per step, and a 3-tap vertical filter producing two output
// rows costs 4 multiplies per (dx, in, out) instead of 6:
//     [d0 d1 d2 d3] -> T0 = d0 - d2, T1 = d1 + d2, T2 = d2 - d1, T3 = d1 - d3            (input transform, owner lane)
//     m_j = sum over (dx, in) of  U_j[dx][out][in] * T_j(column + dx - 1)                  (U = G g, folded on the host)
//     y0 = m0 + m1 + m2,  y1 = m1 - m2 - m3                                                (output transform)
// The owner lane transforms its own column and publishes T (4 values per channel and row pair instead of 2 rows), the
// neighbours read it through the row rings as before.  The packed FFMA2 pair runs over the TRANSFORM index -- (m0, m1) and
// (m2, m3) of one output accumulate over the input channels one after the other -- so there are no (even, odd) partial sums
// to close: y0 = m0 + (m1 + b) + m2 and y1 = (m1 + b) - m2 - m3 cost 5 adds per output and row pair.  Per row: 88 instead of
// 124 FFMA2, 34 instead of 51 LDCU.128 (the transforms amortise badly over 4 channels); 12.6 instead of 10.5 M patches/s
// data -> latent, 11.8 instead of 9.4 M latent -> data (two 8-warp CTAs per SM there, launch_chain_wino).
//
// Reference semantics: identical to nf_kernels.cu / nf_coupling.cuh (layers.py:117-130, 333-375, 452-498, 555-583,
// 651-674; noise_flow_model.py:394-480); results differ from the direct form by fp32 rounding only.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "nf_params.h"
#include "nf_kernels.h"
#include "nf_rng.cuh"
#include "nf_chain_dev.cuh"

#if !NF_Z_IN_TMEM
#error "nf_wino.cu keeps the resident patches in tensor memory"
#endif

namespace nf {
namespace wino {

// NfCouplingP with the 3x3 filters in the transformed domain: index j = 0..3 instead of dy = 0..2
struct alignas(16) CouplingW {
    float a[4][4];
    float ainv[4][4];
    float w1[3][4][2][4];   // [dx][o][i][j]: the four transformed taps of one (dx, out, in) are one LDCU.128
    float w2[4][4][2];      // [o][i] twice: the 1x1 conv runs packed over the two rows of a step
    float w3[3][4][4][4];   // [dx][o][i][j]
    float b1[4];
    float b2[4];
    float b3[3][3][4];
    float scale;
    int32_t has_mix;
    float pad_[2];
};
struct ModelParamsW {
    int32_t n_layers;
    int32_t n_rows;
    int32_t pad_[2];
    int32_t op[NF_MAX_LAYERS];
    int32_t slot[NF_MAX_LAYERS];
    CouplingW cp[NF_MAX_COUPLINGS];
    NfMixP mix[NF_MAX_MIX];
    NfScaleP sc[NF_MAX_SCALE];
};
static_assert(sizeof(ModelParamsW) <= 32 * 1024 - 256, "kernel parameter space");

struct __align__(16) WarpSmemW {
    float4 hw[1][4][34];   // conv-3 input tile: [channel][column + 1] = (T0, T1, T2, T3) of that h2 channel; [0], [33] = zero halo
    float4 xw[1][2][34];   // conv-1 input tile: [channel][column + 1] = (T0, T1, T2, T3) of that x0 channel
};

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(float4 v) { return make_float2(v.z, v.w); }

__device__ __forceinline__ void ld_rows(const ZStore& zs, int r, float4& a, float4& b) {   // rows r, r + 1 (r even)
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(zs.taddr + (uint32_t)(r * 4)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]));
    a = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
    b = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
}
__device__ __forceinline__ void st_rows(const ZStore& zs, int r, const float4& a, const float4& b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(zs.taddr + (uint32_t)(r * 4)), "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
                    "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)), "r"(__float_as_uint(b.z)), "r"(__float_as_uint(b.w)) : "memory");
}

// affine update of one pixel (stage C tail).  With r = 1 / (exp(2 h) + 1):  tanh(h) = 1 - 2 r, so the exponent of
// exp(+-log_scale) = exp2(+-s2 * tanh(h)), s2 = rescaling_scale * log2(e), is ONE FFMA of r, and the log-det contribution
// scale * (tanh(h2) + tanh(h3)) is accumulated as sum(r) and closed once per pass (wino_pass): 8 instead of 11 instructions per
// channel.  shift = h3[0:2], log_scale = scale * tanh(h3[2:4])                                       (layers.py:362 / :342)
template <bool INV>
__device__ __forceinline__ void affine(const float (&h3)[4], const float s2, float4& z, float& rsum, const bool has_mix, const float2 (&am)[4][2]) {
    const float r0 = __fdividef(1.f, exp2f(h3[2]) + 1.f);                    // h3[2:4] arrive times 2 log2(e): to_winograd folds it
    const float r1 = __fdividef(1.f, exp2f(h3[3]) + 1.f);                    // into conv-3's filters and bias
    rsum += r0 + r1;
    if (INV) {
        z.z = fmaf(z.z, exp2f(fmaf(-2.f * s2, r0, s2)), h3[0]);              // layers.py:363-367
        z.w = fmaf(z.w, exp2f(fmaf(-2.f * s2, r1, s2)), h3[1]);
    } else {
        z.z = (z.z - h3[0]) * exp2f(fmaf(2.f * s2, r0, -s2));                // layers.py:343-347
        z.w = (z.w - h3[1]) * exp2f(fmaf(2.f * s2, r1, -s2));
        if (has_mix) z = mix4r(z, am);                                       // Conv2d1x1._forward, layers.py:113-114
    }
}

// Pair step u = 0..16.  The three stages of a step hand their tile to the NEXT STAGE OF THE SAME STEP: two __syncwarp per
// step, single-buffered tiles, 3 guarded steps per pass (u = 0, 1, 16), and the rows stage A mixes stay in registers until stage C
// of the next step finishes them (one tcgen05.ld.x8 and one st.x8 per step).  (A first version passed tiles to the NEXT step -- one __syncwarp, stages free
// to interleave, but 19 steps of which 7 guarded: 11.56 instead of 11.92 M patches/s.)
//   stage A (rows 2u, 2u+1)  : z <- z.A (inverse only); with the retained x0 rows 2u-2, 2u-1 publish the conv-1 input tile
//                              whose outputs are h1 rows (2u-1, 2u)
//   stage B: conv-1 rows (2u-1, 2u) -> BN+ReLU, 1x1 conv, BN+ReLU -> h2 rows; with the retained h2 rows 2u-3, 2u-2 publish the
//            conv-3 input tile whose outputs are rows (2u-2, 2u-1)
//   stage C: conv-3 rows (2u-2, 2u-1) + edge-indicator bias -> tanh/exp affine update of z + log-det
template <bool INV, bool GUARDED, class CP>
__device__ __forceinline__ void wino_step(const CP& P, WarpSmemW& s, const ZStore& zs, const int lane, const int u, const bool has_mix,
                                          float2 (&xp)[2], float4 (&hp)[2], float4 (&zp)[2], float& rsum, const float s2, const float2 (&am)[4][2], const float (&b3m)[4]) {
    const float2 zero2 = make_float2(0.f, 0.f);
    const bool do_a = !GUARDED || u <= 15;
    const bool do_c = !GUARDED || u >= 1;
    // Rows (2u, 2u+1) come from tensor memory; rows (2u-2, 2u-1), which stage C finishes, are the (mixed) rows stage A of the
    // previous step left in registers: a row is read once and written once per coupling.
    zs.commit();
    float4 za0 = make_float4(0.f, 0.f, 0.f, 0.f), za1 = za0, zc0 = zp[0], zc1 = zp[1];
    if (do_a) ld_rows(zs, 2 * u, za0, za1);
    // ---------------- stage A
    float4 tx[2], th[4];
    {
        float2 d2 = zero2, d3 = zero2;
        if (do_a) {
            if (INV && has_mix) {
                za0 = mix4r(za0, am);                           // Conv2d1x1._inverse, layers.py:117-119
                za1 = mix4r(za1, am);
            }
            d2 = lo2(za0);
            d3 = lo2(za1);
        }
        // channel c: (T0, T1, T2, T3) = (d0 - d2, d1 + d2, d2 - d1, d1 - d3)
        tx[0] = make_float4(xp[0].x - d2.x, xp[1].x + d2.x, d2.x - xp[1].x, xp[1].x - d3.x);
        tx[1] = make_float4(xp[0].y - d2.y, xp[1].y + d2.y, d2.y - xp[1].y, xp[1].y - d3.y);
        s.xw[0][0][lane + 1] = tx[0];
        s.xw[0][1][lane + 1] = tx[1];
        xp[0] = d2;
        xp[1] = d3;
    }
    // ---------------- stage B.  The centre tap is this lane's own tile (still in registers): its FFMA2 run BEFORE the
    // __syncwarp that publishes the tile to the neighbours and cover the shared-memory round trip of the other two taps.
    {
        float2 m01[4], m23[4];
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                m01[o] = ffma2(lo2(tx[c]), ld2(&P.w1[1][o][c][0]), c == 0 ? zero2 : m01[o]);
                m23[o] = ffma2(hi2(tx[c]), ld2(&P.w1[1][o][c][2]), c == 0 ? zero2 : m23[o]);
            }
        __syncwarp();
#pragma unroll
        for (int dx = 0; dx < 3; dx += 2) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float4 t = s.xw[0][c][lane + dx];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    m01[o] = ffma2(lo2(t), ld2(&P.w1[dx][o][c][0]), m01[o]);
                    m23[o] = ffma2(hi2(t), ld2(&P.w1[dx][o][c][2]), m23[o]);
                }
            }
        }
        float ha[4], hb[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const float sb = m01[o].y + P.b1[o];                                           // m1 + bias: in both rows with a plus
            ha[o] = fmaxf((m01[o].x + sb) + m23[o].x, 0.f);                                // y0 = m0 + m1 + m2; BN folded, ReLU
            hb[o] = fmaxf((sb - m23[o].x) - m23[o].y, 0.f);                                // y1 = m1 - m2 - m3
        }
        float ea[4], eb[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {   // 1x1 conv, packed over the two rows: channel c of (row a, row b) times (w, w)
            float2 acc = ffma2(make_float2(ha[0], hb[0]), ld2(&P.w2[o][0][0]), zero2);
#pragma unroll
            for (int c = 1; c < 4; ++c) acc = ffma2(make_float2(ha[c], hb[c]), ld2(&P.w2[o][c][0]), acc);
            ea[o] = fmaxf(acc.x + P.b2[o], 0.f);
            eb[o] = fmaxf(acc.y + P.b2[o], 0.f);
        }
        float4 e2 = make_float4(ea[0], ea[1], ea[2], ea[3]), e3 = make_float4(eb[0], eb[1], eb[2], eb[3]);
        if (GUARDED && u == 0) e2 = make_float4(0.f, 0.f, 0.f, 0.f);    // h2 row -1: SAME padding of conv-3
        if (GUARDED && u == 16) e3 = make_float4(0.f, 0.f, 0.f, 0.f);   // h2 row 32
        th[0] = make_float4(hp[0].x - e2.x, hp[1].x + e2.x, e2.x - hp[1].x, hp[1].x - e3.x);
        th[1] = make_float4(hp[0].y - e2.y, hp[1].y + e2.y, e2.y - hp[1].y, hp[1].y - e3.y);
        th[2] = make_float4(hp[0].z - e2.z, hp[1].z + e2.z, e2.z - hp[1].z, hp[1].z - e3.z);
        th[3] = make_float4(hp[0].w - e2.w, hp[1].w + e2.w, e2.w - hp[1].w, hp[1].w - e3.w);
#pragma unroll
        for (int c = 0; c < 4; ++c) s.hw[0][c][lane + 1] = th[c];
        hp[0] = e2;
        hp[1] = e3;
    }
    // ---------------- stage C (centre tap from registers before the __syncwarp, as in stage B)
    float2 m01[4], m23[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            m01[o] = ffma2(lo2(th[c]), ld2(&P.w3[1][o][c][0]), c == 0 ? zero2 : m01[o]);
            m23[o] = ffma2(hi2(th[c]), ld2(&P.w3[1][o][c][2]), c == 0 ? zero2 : m23[o]);
        }
    __syncwarp();
    if (do_c) {
#pragma unroll
        for (int dx = 0; dx < 3; dx += 2) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 t = s.hw[0][c][lane + dx];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    m01[o] = ffma2(lo2(t), ld2(&P.w3[dx][o][c][0]), m01[o]);
                    m23[o] = ffma2(hi2(t), ld2(&P.w3[dx][o][c][2]), m23[o]);
                }
            }
        }
        float h3a[4], h3b[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float ba = b3m[o], bb = b3m[o];
            if (GUARDED) {   // row class of the edge-indicator bias (rows 0 and 31)
                const int cc = lane == 0 ? 0 : (lane == 31 ? 2 : 1);
                if (u == 1) ba = P.b3[0][cc][o];
                if (u == 16) bb = P.b3[2][cc][o];
            }
            h3a[o] = (m01[o].x + (m01[o].y + ba)) + m23[o].x;      // y0 = m0 + m1 + m2
            h3b[o] = ((m01[o].y + bb) - m23[o].x) - m23[o].y;      // y1 = m1 - m2 - m3
        }
        affine<INV>(h3a, s2, zc0, rsum, has_mix, am);
        affine<INV>(h3b, s2, zc1, rsum, has_mix, am);
    }
    if (do_c) st_rows(zs, 2 * u - 2, zc0, zc1);
    zp[0] = za0;
    zp[1] = za1;
}

template <bool INV, class CP>
__device__ __forceinline__ void wino_pass(const CP& P, WarpSmemW& s, const ZStore& zs, const int lane, float& ldj) {
    const bool has_mix = P.has_mix != 0;
    float2 am[4][2];
    load_mix_regs<INV>(P, am, s.xw[0][0][0].x);
    float b3m[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) b3m[o] = lane == 0 ? P.b3[1][0][o] : (lane == 31 ? P.b3[1][2][o] : P.b3[1][1][o]);
    float2 xp[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float4 hp[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    float4 zp[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    const float s2 = P.scale * 1.4426950408889634f;
    float rsum = 0.f;
#pragma unroll 1
    for (int u = 0; u < 17; ++u) {
        if (u >= 2 && u <= 15) wino_step<INV, false>(P, s, zs, lane, u, has_mix, xp, hp, zp, rsum, s2, am, b3m);
        else                   wino_step<INV, true>(P, s, zs, lane, u, has_mix, xp, hp, zp, rsum, s2, am, b3m);
    }
    __syncwarp();
    // log-det of the pass: this lane saw 32 rows x 2 channels;  sum tanh = 64 - 2 sum r       (layers.py:372 / :352)
    const float lsum = P.scale * fmaf(-2.f, rsum, 64.f);
    ldj += INV ? lsum : -lsum;
}

#define NFW_FAST_SLOTS 8
template <bool INV>
__device__ __forceinline__ void wino_dispatch(const ModelParamsW& mp, WarpSmemW& s, const ZStore& zs, int lane, float& ldj, int slot) {
    switch (slot) {
#define NFW_CASE(K) case K: wino_pass<INV>(mp.cp[K], s, zs, lane, ldj); break;
        NFW_CASE(0) NFW_CASE(1) NFW_CASE(2) NFW_CASE(3) NFW_CASE(4) NFW_CASE(5) NFW_CASE(6) NFW_CASE(7)
#undef NFW_CASE
        default: wino_pass<INV>(mp.cp[slot], s, zs, lane, ldj); break;
    }
}

template <bool INV, int TMEM_COLS>
__global__ void __launch_bounds__(NF_MAX_CTA_THREADS, 1)
nf_chain_wino_kernel(const __grid_constant__ ModelParamsW mp, const NfChainArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    WarpSmemW& s = reinterpret_cast<WarpSmemW*>(smem_raw)[warp];

    __shared__ uint32_t tmem_base_smem;
    // 128 tensor-memory columns per group of four warps (a power of two): 16 warps take the whole tensor memory of the SM
    // (512 columns x 128 lanes = 16 resident patches); two 8-warp CTAs per SM take half each (launch_chain_wino)
    // (TMEM_COLS is a template parameter: as a run-time value it flips ptxas' register allocation of the coupling step)
    if (warp == 0) {
        const uint32_t tmem_cols = TMEM_COLS;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_smem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const ZStore zs = {tmem_base_smem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128)};
    for (int e = lane; e < (int)(sizeof(WarpSmemW) / 16); e += 32) reinterpret_cast<float4*>(&s)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();

    const long long stride = (long long)gridDim.x * warps_per_cta;
    for (long long base = (long long)blockIdx.x * warps_per_cta; base < a.n; base += stride) {
        const bool active = base + warp < a.n;
        const long long p = active ? base + warp : a.n - 1;
        int row = a.rows ? a.rows[p] : a.default_row;
        row = min(max(row, 0), NF_MAX_ROWS - 1);
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
        const int l0 = INV ? a.first_layer : a.last_layer - 1, l1 = INV ? a.last_layer : a.first_layer - 1, dl = INV ? 1 : -1;
        // The 16 warps of a CTA move through the chain in lock step (the __syncthreads below), so they would all wait on HBM
        // at the same time: in the prologue (in: 16 KB per patch) and at the SDN layer (y: 16 KB).  Two layers (~45 us) before
        // either is read, lane 0 asks for it as one bulk L2 prefetch: the reads then hit L2.  Footprint in L2: <= 32 KB per warp.
        for (int l = l0; l != l1; l += dl) {
            const int op = mp.op[l], slot = mp.slot[l];
            switch (op) {
                case NF_KOP_COUPLING: wino_dispatch<INV>(mp, s, zs, lane, ldj, slot); break;
                case NF_KOP_MIX: mix_pass<INV>(mp.mix[slot], zs); break;
                case NF_KOP_SDN:
                    sdn_pass<INV>(reinterpret_cast<const float4*>(a.y) + p * NF_PIXELS, mp.sc[slot].t[row][0], mp.sc[slot].t[row][1], zs, lane, ldj);
                    break;
                case NF_KOP_GAIN:
                    gain_pass<INV>(mp.sc[slot].t[row][0], mp.sc[slot].t[row][1], mp.sc[slot].t[row][2], zs, lane, ldj);
                    break;
                default: break;
            }
            __syncthreads();   // layer boundary: keeps the CTA's warps in the same loop body (I-cache)
        }

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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        const uint32_t tmem_cols = TMEM_COLS;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base_smem), "r"(tmem_cols) : "memory");
    }
}

// host: G g of the vertical taps, in double
static void to_winograd(const NfModelParams& mp, ModelParamsW& w) {
    w.n_layers = mp.n_layers;
    w.n_rows = mp.n_rows;
    w.pad_[0] = w.pad_[1] = 0;
    for (int l = 0; l < NF_MAX_LAYERS; ++l) { w.op[l] = mp.op[l]; w.slot[l] = mp.slot[l]; }
    for (int k = 0; k < NF_MAX_MIX; ++k) w.mix[k] = mp.mix[k];
    for (int k = 0; k < NF_MAX_SCALE; ++k) w.sc[k] = mp.sc[k];
    for (int k = 0; k < NF_MAX_COUPLINGS; ++k) {
        const NfCouplingP& p = mp.cp[k];
        CouplingW& q = w.cp[k];
        for (int i = 0; i < 16; ++i) { (&q.a[0][0])[i] = (&p.a[0][0])[i]; (&q.ainv[0][0])[i] = (&p.ainv[0][0])[i]; (&q.w2[0][0][0])[2 * i] = (&q.w2[0][0][0])[2 * i + 1] = (&p.w2[0][0])[i]; }
        for (int i = 0; i < 4; ++i) { q.b1[i] = p.b1[i]; q.b2[i] = p.b2[i]; }
        // conv-3's outputs 2, 3 only feed tanh: they leave the convolution times 2 log2(e), the argument exp2 wants (affine())
        const double k23 = 2.8853900817779268;
        for (int i = 0; i < 36; ++i) (&q.b3[0][0][0])[i] = (float)(((i & 3) >= 2 ? k23 : 1.0) * (double)(&p.b3[0][0][0])[i]);
        q.scale = p.scale;
        q.has_mix = p.has_mix;
        q.pad_[0] = q.pad_[1] = 0.f;
        for (int dx = 0; dx < 3; ++dx)
            for (int o = 0; o < 4; ++o) {
                for (int i = 0; i < 2; ++i) {
                    const double g0 = p.w1[0][dx][o][i], g1 = p.w1[1][dx][o][i], g2 = p.w1[2][dx][o][i];
                    q.w1[dx][o][i][0] = (float)g0;
                    q.w1[dx][o][i][1] = (float)(0.5 * (g0 + g1 + g2));
                    q.w1[dx][o][i][2] = (float)(0.5 * (g0 - g1 + g2));
                    q.w1[dx][o][i][3] = (float)g2;
                }
                for (int i = 0; i < 4; ++i) {
                    const double ko = o >= 2 ? k23 : 1.0;
                    const double g0 = ko * p.w3[0][dx][o][i], g1 = ko * p.w3[1][dx][o][i], g2 = ko * p.w3[2][dx][o][i];
                    q.w3[dx][o][i][0] = (float)g0;
                    q.w3[dx][o][i][1] = (float)(0.5 * (g0 + g1 + g2));
                    q.w3[dx][o][i][2] = (float)(0.5 * (g0 - g1 + g2));
                    q.w3[dx][o][i][3] = (float)g2;
                }
            }
    }
}

}  // namespace wino

bool wino_program_supported(const NfModelParams& mp, const NfChainArgs& a) {
    if (a.bn_stage != 0) return false;   // batch-statistics probes run on the direct-form kernel
    for (int l = a.first_layer; l < a.last_layer; ++l)
        if (mp.op[l] == NF_KOP_COUPLING) return true;
    return false;
}

cudaError_t launch_chain_wino(const NfModelParams& mp, const NfChainArgs& args, bool inverse, int num_sms, cudaStream_t stream) {
    if (args.n <= 0) return cudaSuccess;
    wino::ModelParamsW w;
    wino::to_winograd(mp, w);
    // CTA shape.  One 16-warp CTA per SM owns the whole tensor memory; two 8-warp CTAs take half each, the layer lock step then
    // binds 8 warps and the two halves of an SM drift apart.  Measured at 65 536 patches (profiles/r06_ab_cta_warps.log):
    // sampling with in-kernel Philox 10.76 -> 11.79 M patches/s with two CTAs (one half's FMA-bound coupling passes cover the
    // other half's ALU-bound Philox prologue), log_prob 12.58 -> 12.44 M.  So: two CTAs when the kernel draws its own noise,
    // one otherwise; NF_WINO_CTA_WARPS=8 / 16 forces either (A/B switch).
    static const int forced_warps = [] { const char* e = getenv("NF_WINO_CTA_WARPS"); const int v = e ? atoi(e) : 0; return (v == 8 || v == 16) ? v : 0; }();
    const int cta_warps = forced_warps ? forced_warps : ((!inverse && args.in == nullptr) ? 8 : NF_MAX_WARPS_PER_CTA);
    num_sms *= NF_MAX_WARPS_PER_CTA / cta_warps;   // resident CTAs
    int warps = cta_warps;
    {   // CTA shape against wave quantisation, as launch_chain
        const long long g = args.n < (long long)num_sms ? args.n : (long long)num_sms;
        const long long per_sm = (args.n + g - 1) / g, rounds = (per_sm + warps - 1) / warps;
        warps = (int)((per_sm + rounds - 1) / rounds);
    }
    const size_t smem = (size_t)warps * sizeof(wino::WarpSmemW);
    // CTAs of more than 8 warps own the whole tensor memory (512 columns); smaller ones take 256, so two fit on an SM
    typedef void (*kern_t)(const wino::ModelParamsW, const NfChainArgs);
    const int half = cta_warps <= 8 ? 1 : 0;
    const kern_t kern = inverse ? (half ? wino::nf_chain_wino_kernel<true, 256> : wino::nf_chain_wino_kernel<true, 512>)
                                : (half ? wino::nf_chain_wino_kernel<false, 256> : wino::nf_chain_wino_kernel<false, 512>);
    static bool attr_done[NF_MAX_DEVICES][2][2] = {};
    const int dev = device_slot(), k = inverse ? 0 : 1;
    if (!attr_done[dev][k][half]) {
        const int max_smem = NF_MAX_WARPS_PER_CTA * (int)sizeof(wino::WarpSmemW);
        const cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e != cudaSuccess) return e;
        attr_done[dev][k][half] = true;
    }
    long long ctas = (args.n + warps - 1) / warps;
    if (ctas > num_sms) ctas = num_sms;
    kern<<<(unsigned)ctas, warps * 32, smem, stream>>>(w, args);
    return cudaGetLastError();
}

}  // namespace nf
