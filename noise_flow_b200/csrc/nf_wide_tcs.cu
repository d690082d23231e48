// Wide coupling nets, widths 256 / 512 (the reference's default `--width 512`, sidd/ArgParser.py:43) on the tensor cores
// with STREAMED weights.  Same formulation as nf_wide_tc.cu -- TMEM lane = pixel, activations go accumulator -> ReLU ->
// bf16 (hi, lo) -> A operand without leaving tensor memory, conv-3 as a 1x1 GEMM + shifted sum -- but at these widths
// neither the weights of a coupling (352 KB / 1.2 MB) fit shared memory nor a tile's activations + accumulators the 512
// TMEM columns, so a 128-pixel tile is walked in chunks:
//
//   for pass p (256 conv-2 output channels):                          TMEM columns: A1 32 | D1c 64 | A2c/A3c 64 | D2 256 | D3 96
//     for kc (64 hidden channels):   conv-1 chunk  D1c = A1 . B1[kc]                      4 MMAs, N = 64
//                                    ReLU / split  A2c
//                                    conv-2 part   D2 += A2c . B2[p][kc]                  12 MMAs, N = 256
//     bias MMA, then for nc (64 of the pass's channels):  ReLU / split A3c,  D3 += A3c . B3[p][nc]   8 MMAs, N = 96 / 48
//
// Weight blocks (72 KB: B1 chunk + B2 chunk; 56 KB: bias + B3 of a pass) travel global/L2 -> shared memory through a
// two-slot ring filled by the TMA engine (cp.async.bulk + mbarrier complete_tx) one block ahead of the MMAs; a slot is
// handed back by tcgen05.commit when the MMAs that read it have retired.  One group of 512 threads per CTA (the four warps
// of a TMEM lane quarter split a chunk's 64 channels), one patch at a time.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_kernels.h"
#include "nf_params.h"
#include "nf_rng.cuh"
#include "nf_wide.h"
#include "nf_wide_tc_common.cuh"

namespace nf {
namespace wtc {

template <int W>
struct SCfg {
    static_assert(W == 256 || W == 512, "streamed-weights kernel: width 256 / 512");
    static constexpr int P = W / 256;            // conv-2 output passes
    static constexpr int NK = W / 64;            // 64-channel chunks of the hidden layer
    static constexpr int GT = COMPUTE_THREADS;   // one group
    // TMEM columns
    static constexpr uint32_t C_A1 = 0, C_D1 = 32, C_A2 = 96, C_D2 = 160, C_D3 = 416;
    static constexpr uint32_t B1C = 8192, B2C = 65536, BB2 = 8192, B3P = 49152;
    static constexpr uint32_t SLOT = B1C + B2C;
};

struct __align__(128) SSmem {
    unsigned char slot[2][73728];
    float4 z[NF_PIXELS];
    float4 pre[NF_PIXELS];
    float4 ex[2][4][32];
    float hdr[128];
    float red[64];
    float sacc[1024];
    uint64_t full[2], empty[2], mbar;
    uint32_t tmem_base;
    uint32_t pad_[5];
};

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// accumulator columns [d, d + 16) of this thread's pixel -> ReLU -> (hi, lo) split -> 8 + 8 A-operand columns
__device__ __forceinline__ void relu_split_store16(uint32_t d, uint32_t a_hi, uint32_t a_lo) {
    uint32_t r[16];
    tmem_ld16(d, r);
    tmem_wait_ld();
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) relu_split2(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]), hi[k], lo[k]);
    tmem_st8(a_hi, hi);
    tmem_st8(a_lo, lo);
    tmem_wait_st();
}
// batch-statistics probe over 16 accumulator columns: lanes 0..15 add channel `lane` of this warp's 32 pixels to sacc
__device__ __forceinline__ void probe16(uint32_t d, int lane, float* s_sum, float* s_sq) {
    uint32_t r[16];
    tmem_ld16(d, r);
    tmem_wait_ld();
    float v[16], q[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { v[k] = __uint_as_float(r[k]); q[k] = v[k] * v[k]; }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < off; ++k) {
            const float sv = up ? v[k] : v[k + off], mv = up ? v[k + off] : v[k];
            v[k] = mv + __shfl_xor_sync(0xffffffffu, sv, off);
            const float sq = up ? q[k] : q[k + off], mq = up ? q[k + off] : q[k];
            q[k] = mq + __shfl_xor_sync(0xffffffffu, sq, off);
        }
    }
    const float ts = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16), tq = q[0] + __shfl_xor_sync(0xffffffffu, q[0], 16);
    if (lane < 16) { atomicAdd(s_sum + lane, ts); atomicAdd(s_sq + lane, tq); }
}

// ---- the weight-block schedule of one coupling execution, shared by producer and consumer -------------------------
//   STAGE 0: per tile, per pass: NK blocks A (B1 chunk + B2 chunk), one block B (bias + B3 of the pass)
//   STAGE 1: per tile: NK blocks A' (B1 chunk only)              STAGE 2: per tile, per pass: NK blocks A, one block B' (bias)
template <int W>
__device__ __forceinline__ void produce_coupling(SSmem& S, const CUtensorMap* tmap, int row0, int stage, uint32_t& blk) {
    using C = SCfg<W>;
    using L = NfWideTcLayout;
    const int passes = stage == 1 ? 1 : C::P;
    for (int t = 0; t < 8; ++t)
        for (int p = 0; p < passes; ++p) {
            for (int kc = 0; kc <= C::NK; ++kc) {
                if (kc == C::NK && stage == 1) break;
                const uint32_t s = blk & 1u, use = blk >> 1;
                if (use > 0) mbar_wait(smem_u32(&S.empty[s]), (use - 1u) & 1u);
                const uint32_t fb = smem_u32(&S.full[s]), dst = smem_u32(&S.slot[s][0]);
                if (kc < C::NK) {
                    const bool with_b2 = stage != 1;
                    mbar_arrive_expect_tx(fb, C::B1C + (with_b2 ? C::B2C : 0u));
                    // TMA tensor copies in boxes of 32 rows x 256 B of the blob's 2-D view
                    tma_load_rows(dst, tmap, row0 + (int)((L::off_b1() + kc * C::B1C) / 256u), fb);
                    if (with_b2) {
                        const int r2 = row0 + (int)((L::off_b2(W) + ((size_t)p * C::NK + kc) * C::B2C) / 256u);
                        for (uint32_t o = 0; o < C::B2C; o += 8192u) tma_load_rows(dst + C::B1C + o, tmap, r2 + (int)(o / 256u), fb);
                    }
                } else {
                    const bool with_b3 = stage == 0;
                    mbar_arrive_expect_tx(fb, C::BB2 + (with_b3 ? C::B3P : 0u));
                    tma_load_rows(dst, tmap, row0 + (int)((L::off_bb2(W) + (size_t)p * C::BB2) / 256u), fb);
                    if (with_b3) {
                        const int r3 = row0 + (int)((L::off_b3(W) + (size_t)p * C::B3P) / 256u);
                        for (uint32_t o = 0; o < C::B3P; o += 8192u) tma_load_rows(dst + C::BB2 + o, tmap, r3 + (int)(o / 256u), fb);
                    }
                }
                ++blk;
            }
        }
}

template <int W, bool INV, int STAGE>
__device__ __forceinline__ void tcs_coupling(const float* __restrict__ cblob, SSmem& S, const int i, const uint32_t tmem, uint32_t& mphase,
                                             uint32_t& blk, float& ldj) {
    using C = SCfg<W>;
    using L = NfWideTcLayout;
    constexpr int GT = C::GT;
    const int lane = i & 31, wq = (i >> 5) & 3, h = i >> 7;
    if (i < 128) S.hdr[i] = __ldg(cblob + i);
    group_barrier(0, GT);
    const float* hdr = S.hdr;
    const bool has_mix = hdr[L::H_META] != 0.f;
#pragma unroll 2
    for (int k = 0; k < NF_PIXELS / GT; ++k) {
        const int px = k * GT + i, r = px >> 5, c = px & 31;
        if (INV && has_mix && STAGE == 0) S.z[px] = mix4(S.z[px], hdr + L::H_A);
        const int rc = r == 0 ? 0 : (r == 31 ? 2 : 1), cc = c == 0 ? 0 : (c == 31 ? 2 : 1);
        S.pre[px] = *reinterpret_cast<const float4*>(hdr + L::H_B3 + (rc * 3 + cc) * 4);
    }
    group_barrier(0, GT);
    const bool mix_on_the_fly = STAGE != 0 && INV && has_mix;
    const uint32_t mbar = smem_u32(&S.mbar);
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
    const uint32_t tA1 = tmem + C::C_A1, tD1 = tmem + C::C_D1, tA2 = tmem + C::C_A2, tD2 = tmem + C::C_D2, tD3 = tmem + C::C_D3;
    constexpr int PASSES = STAGE == 1 ? 1 : C::P;

#pragma unroll 1
    for (int t = 0; t < 8; ++t) {
        const int r = 4 * t + wq;
        if (h == 0) {   // A1 (as nf_wide_tc.cu): [x_hi 18 | x_lo 18 | x_hi 18 | 1 1 | 0], kept for the whole tile
            uint32_t a1[32];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int rr = r + dy - 1, cc = lane + dx - 1;
                    float2 x0 = make_float2(0.f, 0.f);
                    if (rr >= 0 && rr <= 31 && cc >= 0 && cc <= 31) {
                        if (mix_on_the_fly) { const float4 zz = mix4(S.z[rr * 32 + cc], hdr + L::H_A); x0 = make_float2(zz.x, zz.y); }
                        else x0 = *reinterpret_cast<const float2*>(&S.z[rr * 32 + cc]);
                    }
                    uint32_t hi, lo;
                    split2(x0.x, x0.y, hi, lo);
                    a1[dy * 3 + dx] = hi;
                    a1[9 + dy * 3 + dx] = lo;
                    a1[18 + dy * 3 + dx] = hi;
                }
            a1[27] = 0x3F803F80u;
            a1[28] = a1[29] = a1[30] = a1[31] = 0u;
            uint32_t lo16[16], hi16[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) { lo16[k] = a1[k]; hi16[k] = a1[16 + k]; }
            tmem_st16(tA1 + lane_sel, lo16);
            tmem_st16(tA1 + lane_sel + 16u, hi16);
            tmem_wait_st();
        }
        tc_fence_before();
        group_barrier(0, GT);
#pragma unroll 1
        for (int p = 0; p < PASSES; ++p) {
#pragma unroll 1
            for (int kc = 0; kc < C::NK; ++kc) {
                const uint32_t s = blk & 1u, use = blk >> 1;
                const uint32_t sb = smem_u32(&S.slot[s][0]);
                // ---------------- conv-1 chunk: D1c[128 x 64] = A1 . B1[kc]
                if (i == 0) {
                    mbar_wait(smem_u32(&S.full[s]), use & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_ts(tD1, tA1 + 8u * j, make_desc(sb + (uint32_t)j * 2u * 1024u, 1024u, 128u), idesc(64), j > 0 ? 1u : 0u);
                    mma_commit(mbar);
                    if (STAGE == 1) mma_commit(smem_u32(&S.empty[s]));
                }
                group_wait_mma(mbar, mphase, 0, i, GT);
                tc_fence_after();
                if (STAGE == 1) {
                    probe16(tD1 + lane_sel + 16u * h, lane, &S.sacc[64 * kc + 16 * h], &S.sacc[W + 64 * kc + 16 * h]);
                    tc_fence_before();
                    group_barrier(0, GT);
                    ++blk;
                    continue;
                }
                // ---------------- ReLU / split -> A2c (this thread: channels 64 kc + 16 h ..)
                relu_split_store16(tD1 + lane_sel + 16u * h, tA2 + lane_sel + 8u * h, tA2 + lane_sel + 32u + 8u * h);
                tc_fence_before();
                group_barrier(0, GT);
                // ---------------- conv-2 partial: D2[128 x 256] += A2c . B2[p][kc]   (hi x hi, lo x hi, hi x lo)
                if (i == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t bk = sb + C::B1C + (uint32_t)q * 2u * 8192u;
                        mma_ts(tD2, tA2 + 8u * q, make_desc(bk, 8192u, 128u), idesc(256), (kc | q) ? 1u : 0u);
                        mma_ts(tD2, tA2 + 32u + 8u * q, make_desc(bk, 8192u, 128u), idesc(256), 1u);
                        mma_ts(tD2, tA2 + 8u * q, make_desc(bk + 4096u, 8192u, 128u), idesc(256), 1u);
                    }
                    mma_commit(smem_u32(&S.empty[s]));      // slot free when these have retired
                    mma_commit(mbar);
                }
                group_wait_mma(mbar, mphase, 0, i, GT);          // A2c / D1c are reused by the next chunk
                tc_fence_after();
                ++blk;
            }
            if (STAGE == 1) continue;
            // ---------------- block B: bias of the pass (through A1's constant-one slots), then conv-3 over the pass's channels
            const uint32_t s = blk & 1u, use = blk >> 1;
            const uint32_t sb = smem_u32(&S.slot[s][0]);
            if (i == 0) {
                mbar_wait(smem_u32(&S.full[s]), use & 1u);
                tc_fence_after();
                mma_ts(tD2, tA1 + 24u, make_desc(sb, 4096u, 128u), idesc(256), 1u);
                mma_commit(mbar);
                if (STAGE == 2) mma_commit(smem_u32(&S.empty[s]));
            }
            group_wait_mma(mbar, mphase, 0, i, GT);
            tc_fence_after();
#pragma unroll 1
            for (int nc = 0; nc < 4; ++nc) {
                if (STAGE == 2) {
                    probe16(tD2 + lane_sel + 64u * nc + 16u * h, lane, &S.sacc[256 * p + 64 * nc + 16 * h], &S.sacc[W + 256 * p + 64 * nc + 16 * h]);
                    continue;
                }
                relu_split_store16(tD2 + lane_sel + 64u * nc + 16u * h, tA2 + lane_sel + 8u * h, tA2 + lane_sel + 32u + 8u * h);
                tc_fence_before();
                group_barrier(0, GT);
                if (i == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t bk = sb + C::BB2 + (uint32_t)(nc * 4 + q) * 2u * 1536u;
                        mma_ts(tD3, tA2 + 8u * q, make_desc(bk, 1536u, 128u), idesc(96), (p | nc | q) ? 1u : 0u);
                        mma_ts(tD3, tA2 + 32u + 8u * q, make_desc(bk, 1536u, 128u), idesc(48), 1u);
                    }
                    if (nc == 3) mma_commit(smem_u32(&S.empty[s]));
                    mma_commit(mbar);
                }
                group_wait_mma(mbar, mphase, 0, i, GT);
                tc_fence_after();
            }
            if (STAGE == 2) { tc_fence_before(); group_barrier(0, GT); }
            ++blk;
        }
        if (STAGE != 0) continue;
        // ---------------- epilogue 3: shifted sum (as nf_wide_tc.cu; the warps h = 0, 1, 2 of a lane quarter take dy = h)
        float s_dy[4] = {0.f, 0.f, 0.f, 0.f};
        if (h < 3) {
            uint32_t a[16], b[16];
            tmem_ld16(tD3 + lane_sel + 16u * h, a);
            tmem_ld16(tD3 + lane_sel + 48u + 16u * h, b);
            tmem_wait_ld();
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const float d0 = __uint_as_float(a[o]) + __uint_as_float(b[o]);
                const float d1 = __uint_as_float(a[4 + o]) + __uint_as_float(b[4 + o]);
                const float d2 = __uint_as_float(a[8 + o]) + __uint_as_float(b[8 + o]);
                float fl = __shfl_up_sync(0xffffffffu, d0, 1);
                float fr = __shfl_down_sync(0xffffffffu, d2, 1);
                if (lane == 0) fl = 0.f;
                if (lane == 31) fr = 0.f;
                s_dy[o] = d1 + (fl + fr);
            }
            if (h == 0) S.ex[0][wq][lane] = make_float4(s_dy[0], s_dy[1], s_dy[2], s_dy[3]);
            if (h == 2) S.ex[1][wq][lane] = make_float4(s_dy[0], s_dy[1], s_dy[2], s_dy[3]);
        }
        tc_fence_before();
        group_barrier(0, GT);
        if (h == 1) {
            float4 acc = make_float4(s_dy[0], s_dy[1], s_dy[2], s_dy[3]);
            if (wq > 0) { const float4 u = S.ex[0][wq - 1][lane]; acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w; }
            if (wq < 3) { const float4 d = S.ex[1][wq + 1][lane]; acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w; }
            float4 pv = S.pre[r * 32 + lane];
            pv.x += acc.x; pv.y += acc.y; pv.z += acc.z; pv.w += acc.w;
            S.pre[r * 32 + lane] = pv;
        }
        if ((h == 0 && wq == 3 && r < 31) || (h == 2 && wq == 0 && r > 0)) {
            const int tr = h == 0 ? r + 1 : r - 1;
            float4 pv = S.pre[tr * 32 + lane];
            pv.x += s_dy[0]; pv.y += s_dy[1]; pv.z += s_dy[2]; pv.w += s_dy[3];
            S.pre[tr * 32 + lane] = pv;
        }
    }
    group_barrier(0, GT);
    if (STAGE != 0) return;
    const float scale = hdr[L::H_META + 1];
#pragma unroll 2
    for (int k = 0; k < NF_PIXELS / GT; ++k) {
        const int px = k * GT + i;
        const float4 pv = S.pre[px];
        float4 z = S.z[px];
        const float ls0 = scale * t_tanh(pv.z), ls1 = scale * t_tanh(pv.w);
        if (INV) {
            z.z = fmaf(z.z, t_exp(ls0), pv.x);
            z.w = fmaf(z.w, t_exp(ls1), pv.y);
            ldj += ls0 + ls1;
        } else {
            z.z = (z.z - pv.x) * t_exp(-ls0);
            z.w = (z.w - pv.y) * t_exp(-ls1);
            ldj -= ls0 + ls1;
            if (has_mix) z = mix4(z, hdr + L::H_AINV);
        }
        S.z[px] = z;
    }
}

template <int W, bool INV>
__global__ void __launch_bounds__(THREADS, 1)
nf_wide_tcs_kernel(const __grid_constant__ CUtensorMap tmap, const NfWideProgram prog, const float* __restrict__ blob, const NfChainArgs a) {
    using C = SCfg<W>;
    constexpr int GT = C::GT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SSmem& S = *reinterpret_cast<SSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool producer = tid >= COMPUTE_THREADS;
    const int i = tid % GT;

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&S.full[b]), 1u); mbar_init(smem_u32(&S.empty[b]), 1u); }
        mbar_init(smem_u32(&S.mbar), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int k = tid; k < 1024; k += THREADS) S.sacc[k] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int n_l = a.last_layer - a.first_layer;

    if (producer) {
        uint32_t blk = 0;
        for (long long p = blockIdx.x; p < a.n; p += gridDim.x)
            for (int step = 0; step < n_l; ++step) {
                const int l = INV ? a.first_layer + step : a.last_layer - 1 - step;
                if (prog.op[l] != NF_KOP_COUPLING) continue;
                const int stage = (a.bn_stage != 0 && step == n_l - 1) ? a.bn_stage : 0;
                if (lane == 0) produce_coupling<W>(S, &tmap, prog.off[l] / NF_TMA_ROW_FLOATS, stage, blk);
                blk = __shfl_sync(0xffffffffu, blk, 0);
            }
    } else {
        const uint32_t tmem = S.tmem_base;
        uint32_t mphase = 0, blk = 0;
        for (long long p = blockIdx.x; p < a.n; p += gridDim.x) {
            int row = a.rows ? a.rows[p] : a.default_row;
            row = min(max(row, 0), NF_MAX_ROWS - 1);
#pragma unroll 2
            for (int k = 0; k < NF_PIXELS / GT; ++k) {
                const int px = k * GT + i;
                float4 v;
                if (a.in) {
                    v = __ldcs(reinterpret_cast<const float4*>(a.in) + p * NF_PIXELS + px);
                    if (!INV) { v.x *= a.temp; v.y *= a.temp; v.z *= a.temp; v.w *= a.temp; }
                } else {
                    v = philox_normal4(a.seed, a.offset, a.patch_base + (unsigned long long)p, (unsigned int)px);
                    v.x *= a.temp; v.y *= a.temp; v.z *= a.temp; v.w *= a.temp;
                }
                S.z[px] = v;
            }
            group_barrier(0, GT);
            float ldj = 0.f;
            for (int step = 0; step < n_l; ++step) {
                const int l = INV ? a.first_layer + step : a.last_layer - 1 - step;
                const int op = prog.op[l];
                const float* pb = blob + prog.off[l];
                if (op == NF_KOP_COUPLING) {
                    const bool probe = a.bn_stage != 0 && step == n_l - 1;
                    if (!probe) tcs_coupling<W, INV, 0>(pb, S, i, tmem, mphase, blk, ldj);
                    else if (a.bn_stage == 1) tcs_coupling<W, INV, 1>(pb, S, i, tmem, mphase, blk, ldj);
                    else tcs_coupling<W, INV, 2>(pb, S, i, tmem, mphase, blk, ldj);
                } else if (op == NF_KOP_MIX) {
                    for (int k = 0; k < NF_PIXELS / GT; ++k) S.z[k * GT + i] = mix4(S.z[k * GT + i], pb + (INV ? 0 : 16));
                } else if (op == NF_KOP_SDN) {
                    const float sa = __ldg(pb + row * 4), sb = __ldg(pb + row * 4 + 1);
                    const float4* yp = reinterpret_cast<const float4*>(a.y) + p * NF_PIXELS;
                    float acc = 0.f;
                    for (int k = 0; k < NF_PIXELS / GT; ++k) {
                        const float4 y = __ldg(yp + k * GT + i);
                        float4 z = S.z[k * GT + i];
                        const float v0 = fmaf(sa, y.x, sb), v1 = fmaf(sa, y.y, sb), v2 = fmaf(sa, y.z, sb), v3 = fmaf(sa, y.w, sb);
                        const float r0 = rsqrtf(v0), r1 = rsqrtf(v1), r2 = rsqrtf(v2), r3 = rsqrtf(v3);
                        if (INV) { z.x *= r0; z.y *= r1; z.z *= r2; z.w *= r3; }
                        else     { z.x *= v0 * r0; z.y *= v1 * r1; z.z *= v2 * r2; z.w *= v3 * r3; }
                        acc += (__logf(v0) + __logf(v1)) + (__logf(v2) + __logf(v3));
                        S.z[k * GT + i] = z;
                    }
                    ldj += INV ? -0.5f * acc : 0.5f * acc;
                } else if (op == NF_KOP_GAIN) {
                    const float mlt = INV ? __ldg(pb + row * 4 + 1) : __ldg(pb + row * 4);
                    for (int k = 0; k < NF_PIXELS / GT; ++k) {
                        float4 z = S.z[k * GT + i];
                        z.x *= mlt; z.y *= mlt; z.z *= mlt; z.w *= mlt;
                        S.z[k * GT + i] = z;
                    }
                    if (i == 0) ldj += INV ? __ldg(pb + row * 4 + 2) : -__ldg(pb + row * 4 + 2);
                }
                group_barrier(0, GT);
            }
            if (a.bn_stage != 0) {
                for (int k = i; k < 2 * W; k += GT) { atomicAdd(a.bn_stats + k, (double)S.sacc[k]); S.sacc[k] = 0.f; }
                group_barrier(0, GT);
                continue;
            }
            float s1 = 0.f, s2 = 0.f;
            float4* dst = a.out ? reinterpret_cast<float4*>(a.out) + p * NF_PIXELS : nullptr;
            for (int k = 0; k < NF_PIXELS / GT; ++k) {
                const float4 z = S.z[k * GT + i];
                if (dst) __stcs(dst + k * GT + i, z);
                s1 += (z.x + z.y) + (z.z + z.w);
                s2 = fmaf(z.x, z.x, fmaf(z.y, z.y, fmaf(z.z, z.z, fmaf(z.w, z.w, s2))));
            }
            ldj = wsum(ldj); s1 = wsum(s1); s2 = wsum(s2);
            if (lane == 0) { S.red[(i >> 5) * 4] = ldj; S.red[(i >> 5) * 4 + 1] = s1; S.red[(i >> 5) * 4 + 2] = s2; }
            group_barrier(0, GT);
            if (i == 0) {
                float t_ldj = 0.f, t1 = 0.f, t2 = 0.f;
                for (int k = 0; k < GT / 32; ++k) { t_ldj += S.red[k * 4]; t1 += S.red[k * 4 + 1]; t2 += S.red[k * 4 + 2]; }
                const float logdet = t_ldj + (INV ? a.ldj_const : -a.ldj_const) + (a.logdet_in ? a.logdet_in[p] : 0.f);
                if (a.logdet) a.logdet[p] = logdet;
                if (a.nll) a.nll[p] = -(logdet - 0.5f * (NF_DIMS * 1.8378770664093453f + t2));
                if (a.sdz) {
                    const float mean = t1 * (1.f / NF_DIMS);
                    a.sdz[p] = sqrtf(fmaxf(t2 * (1.f / NF_DIMS) - mean * mean, 0.f));
                }
            }
            group_barrier(0, GT);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(S.tmem_base), "r"(512u) : "memory");
}

template <int W>
static cudaError_t launch_s(const NfWideProgram& prog, const float* blob, const NfChainArgs& a, bool inverse, int num_sms, cudaStream_t stream) {
    const size_t smem = sizeof(SSmem);
    cudaError_t e = cudaFuncSetAttribute(nf_wide_tcs_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(nf_wide_tcs_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    CUtensorMap tmap;
    e = make_blob_tensor_map(blob, prog.blob_floats, wide_tc_box_rows(W), &tmap);
    if (e != cudaSuccess) return e;
    const long long grid = a.n < (long long)num_sms ? a.n : (long long)num_sms;
    if (inverse) nf_wide_tcs_kernel<W, true><<<(unsigned)grid, THREADS, smem, stream>>>(tmap, prog, blob, a);
    else nf_wide_tcs_kernel<W, false><<<(unsigned)grid, THREADS, smem, stream>>>(tmap, prog, blob, a);
    return cudaGetLastError();
}

}  // namespace wtc

cudaError_t launch_chain_wide_tcs(const NfWideProgram& prog, const float* blob, const NfChainArgs& a, bool inverse, int num_sms,
                                  cudaStream_t stream) {
    if (a.n <= 0) return cudaSuccess;
    switch (prog.width) {
        case 256: return wtc::launch_s<256>(prog, blob, a, inverse, num_sms, stream);
        case 512: return wtc::launch_s<512>(prog, blob, a, inverse, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace nf
