// Fused Noise Flow chain for sm_100a with the coupling net's last convolution on the tensor cores
// ("hybrid" chain kernel).  Same organisation as nf_kernels.cu -- one warp owns one 32x32x4 patch, the patch is
// resident on the SM for the whole chain, a coupling is one software-pipelined pass over the 32 rows -- but conv-3
// (3x3, 4 -> 4, 58 % of the multiply-adds of a coupling) leaves the FP32 pipe:
//
//   * FOUR warps form a group (their 4 x 32 lanes are the 128 TMEM lanes an M = 128 MMA writes); a CTA holds four
//     groups = 16 resident patches.  For image row r every lane writes its pixel's h2 activation, split into fp16
//     (hi, lo) halves, into the group's A tile in shared memory: once as its own pixel's centre tap and once each as the
//     left / right neighbour's tap, so the three horizontal taps of a pixel sit side by side in K and the SAME padding
//     is simply never written.
//   * D[128 x 16] = A[128 x 48] . B[48 x 16]  runs as three tcgen05.mma (kind::f16, fp32 accumulate in TMEM).
//     K = 3 taps x (hi, lo) x 4 channels against W_hi, plus the hi halves again against W_lo, plus three one-hot
//     "column class" slots that carry the conv2d_zeros bias incl. its edge-indicator taps; N = 3 vertical taps x 4
//     outputs.  fp16 (hi, lo) keeps 22 mantissa bits of every activation and weight; the dropped lo x lo term is
//     2^-22 relative.  The four warps of a group take turns at issuing (no issuer warp: a fifth warp on a scheduler
//     would cap everybody at 102 registers): the warp whose turn it is waits on the row's "A tile written" mbarrier.
//   * every warp reads its 12 accumulator columns back two steps later (tcgen05.ld) and adds the three vertical taps
//     into the two pending output rows it carries (8 FADD per pixel instead of 72 FFMA2 + 36 LDCU + 8 FADD).
//
// A tiles and accumulators are double-buffered per group, so the MMAs of row r run while the warps are busy with rows
// r + 1 and r + 2.  Tensor memory: columns 0..383 hold the resident patches of groups 0..2, columns 384..511 the 4 x 2
// accumulator tiles; the patches of group 3 live in shared memory (ZDual).
//
// Reference semantics: identical to nf_kernels.cu / nf_coupling.cuh (layers.py:117-130, 333-375, 452-498, 555-583,
// 651-674; noise_flow_model.py:394-480).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_params.h"
#include "nf_kernels.h"
#include "nf_rng.cuh"
#include "nf_chain_dev.cuh"

#if !NF_Z_IN_TMEM
#error "nf_hybrid.cu keeps the resident patches in tensor memory"
#endif

namespace nf {
namespace hyb {

constexpr int GROUPS = 4;                    // groups of four worker warps per CTA
constexpr int WORKERS = GROUPS * 4;          // resident patches per CTA
constexpr int THREADS = WORKERS * 32;
constexpr int TMEM_GROUPS = 3;               // groups whose patches live in tensor memory; the last group's live in shared memory
constexpr int A_CHUNKS = 6;                  // K = 48 = 6 chunks of 8 fp16 (16 bytes)
constexpr int A_STAGE = A_CHUNKS * 128;      // uint4 per A stage
constexpr uint32_t D_COL0 = 128u * TMEM_GROUPS;   // first accumulator column (384): 4 groups x 2 stages x 16 columns follow

struct __align__(128) Smem {
    uint4 a[GROUPS + 1][2][A_CHUNKS][128];    // A tiles, K-major, no swizzle: [k chunk][row], LBO = 2048 B, SBO = 128 B;
                                              // tile [GROUPS] is a write-only dump for the taps that fall outside the image
    uint4 b[NF_MAX_COUPLINGS][A_CHUNKS][16];  // B tiles per coupling: [k chunk][n], LBO = 256 B, SBO = 128 B
    uint4 z[4][NF_PIXELS];                    // resident patches of the last group: z[row * 32 + lane]
    float2 xr[WORKERS][2][34];                // x0 row ring per worker warp ([0] and [33] = zero halo)
    uint64_t dfull[GROUPS][2];                // accumulators of row r complete (tcgen05.commit)
    uint64_t afull[GROUPS][2];                // A tile of row r written by the group's four warps
    uint32_t tmem_base;
    uint32_t pad_[3];
};
static_assert(sizeof(Smem) <= 227 * 1024, "Smem too large");

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);   // version 1, SWIZZLE_NONE
}
// kind::f16 instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, N = 16, M = 128
constexpr uint32_t IDESC = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count));
}
// try_wait suspends the thread until the phase completes or the time hint (ns) runs out: a long hint costs no latency and
// keeps a waiting warp from spending issue slots on polling (without it: ~9 polls per row step)
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tNFH_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
                 "@p bra NFH_DONE;\n\tbra NFH_WAIT;\n\tNFH_DONE:\n\t}\n" :: "r"(mbar), "r"(parity), "r"(2000u) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_issue_ld4(uint32_t taddr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
}
__device__ __forceinline__ void tmem_issue_ld12(uint32_t taddr, uint32_t (&r)[12]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]) : "r"(taddr + 8u));
}
// tcgen05.wait::ld with every register the pending loads write as an in/out operand of the statement
__device__ __forceinline__ void tmem_wait_ld20(uint32_t (&a)[8], uint32_t (&b)[12]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                   "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]) :: "memory");
}

// Where a worker keeps its resident patch z[32 rows][32 lanes] (float4 per pixel; lane = image column): groups 0..2 in
// TENSOR MEMORY (384 columns, as in nf_kernels.cu), the last group in shared memory -- the accumulator tiles of the four
// groups take the remaining 128 TMEM columns.  `sm` is warp-uniform.  Same interface as ZStore (nf_chain_dev.cuh).
struct ZDual {
    uint32_t taddr;   // (first TMEM lane of this warp << 16) | first column of this warp's patch
    uint4* zsm;       // this lane's column of the shared-memory patch: zsm[r * 32]
    bool sm;
    __device__ __forceinline__ void issue_ld(int r, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) const {
        if (sm) { const uint4 v = zsm[r * 32]; a = v.x; b = v.y; c = v.z; d = v.w; }
        else tmem_issue_ld4(taddr + (uint32_t)(r * 4), a, b, c, d);
    }
    __device__ __forceinline__ float4 load(int r) const {
        uint32_t a, b, c, d;
        issue_ld(r, a, b, c, d);
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d) :: "memory");
        return make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d));
    }
    __device__ __forceinline__ void store(int r, float4 z) const {
        if (sm) zsm[r * 32] = make_uint4(__float_as_uint(z.x), __float_as_uint(z.y), __float_as_uint(z.z), __float_as_uint(z.w));
        else asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                          :: "r"(taddr + (uint32_t)(r * 4)), "r"(__float_as_uint(z.x)), "r"(__float_as_uint(z.y)),
                             "r"(__float_as_uint(z.z)), "r"(__float_as_uint(z.w)) : "memory");
    }
    __device__ __forceinline__ void commit() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
};

// ReLU + fp16 (hi, lo) split of two floats, first value in the low half:  hi = rz(max(v, 0)) -- truncation keeps
// hi <= v, so the remainder of a positive v is never negative -- and lo = rn(max(v - hi, 0)) (a negative v has
// hi = 0 and a negative remainder, which clamps to 0).  hi + lo = max(v, 0) to 2^-22 relative (2^-25 absolute below
// 2^-3, where lo is subnormal); values above 2 x 65504 come out as inf and poison the result visibly.
__device__ __forceinline__ void relu_split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    const float2 r = ffma2(hf, make_float2(-1.f, -1.f), make_float2(v0, v1));
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r.y), "f"(r.x));
}
__device__ __forceinline__ uint16_t half_bits(float v) {
    const __half h = __float2half_rn(v);
    return *reinterpret_cast<const uint16_t*>(&h);
}

// ---- B tile of one coupling: element (k chunk j, column n, slot kk).  n = dy * 4 + o (12 of 16 columns used).
//   chunks 0..2: horizontal tap dx = j, slots [hi c0..c3 | lo c0..c3] of the activation  -> W_hi[dy][dx][o][c] twice
//   chunk  3   : [hi of tap 0 | hi of tap 1]                                              -> W_lo
//   chunk  4   : [hi of tap 2 | unused]                                                   -> W_lo, 0
//   chunk  5   : one-hot column class [left, mid, right, 0] twice                         -> bias (hi | lo), dy = 1 only
__device__ __forceinline__ float b_value(const NfCouplingP& P, int j, int n, int kk) {
    if (n >= 12) return 0.f;
    const int dy = n >> 2, o = n & 3, c = kk & 3;
    const bool second = kk >= 4;
    if (j == 5) {
        if (dy != 1 || c == 3) return 0.f;
        const float w = P.b3[1][c][o], whi = __half2float(__float2half_rn(w));
        return second ? w - whi : whi;
    }
    int dx;
    bool lo_part;
    if (j < 3) { dx = j; lo_part = false; }
    else if (j == 3) { dx = second ? 1 : 0; lo_part = true; }
    else { if (second) return 0.f; dx = 2; lo_part = true; }
    const float w = P.w3[dy][dx][o][c], whi = __half2float(__float2half_rn(w));
    return lo_part ? w - whi : whi;
}

struct Worker {          // per worker-warp constants
    uint4* a_row;        // &S.a[group][0][0][32 * quarter + lane]
    uint4* a_right;      // same entry of the pixel to my right (lane 31: of the dump tile -- the SAME padding stays zero)
    uint4* a_left;       // ... to my left (lane 0: dump tile)
    float2 (*xr)[34];    // this warp's x0 row ring
    uint32_t a_tile;     // shared address of S.a[group][0] (stage 1: + A_STAGE * 16)            } warp-uniform:
    uint32_t afull;      // shared address of S.afull[group][0] (stage 1: + 8)                   } operands of the
    uint32_t dfull;      // shared address of S.dfull[group][0] (stage 1: + 8)                   } MMAs the group's
    uint32_t d_mma;      // TMEM address of the group's accumulator stage 0 (stage 1: + 16)      } last warp issues
    uint32_t d_taddr;    // ... of this warp's quarter of it
    int quarter;         // this warp's index in its group: it issues the MMAs of the rows r with (r & 3) == quarter
};

// One row step of a coupling pass.  Schedule of coupling_step in nf_coupling.cuh with stage C one step later (t = 0..36):
//   stage B (i = t-1): scatter x0 row i into the pending conv-1 rows; the finished row r = t-2 goes through BN+ReLU,
//                      the 1x1 conv, BN+ReLU, is split and written into A tile (r & 1); arrive on afull
//   stage A (t)      : z[t] <- z[t].A (inverse only); publish x0 row t
//   stage C (j = t-4): accumulators of h2 row j (MMAs triggered TWO steps ago, so the wait below never stalls) -> the
//                      three vertical taps go into the pending conv-3 rows; the finished row q = j-1 = t-5 gets the affine
//                      update + log-det
template <bool INV, bool GUARDED, class CP>
__device__ __forceinline__ void hyb_step(const CP& P, const Worker& wk, const ZDual& zs, const uint32_t b_addr, const int lane, const int t,
                                         const bool has_mix, Acc4& b_old, Acc4& b_mid, float (&c_old)[4], float (&c_mid)[4],
                                         float& ldj, const float2 (&am)[4][2], const float2 (&w2c)[4][2]) {
    const float2 zero2 = make_float2(0.f, 0.f);
    const bool do_a = !GUARDED || t < 32;
    const bool b_fma = !GUARDED || (t >= 1 && t <= 32);
    const bool b_emit = !GUARDED || (t >= 2 && t <= 33);
    const bool c_fma = !GUARDED || (t >= 4 && t <= 35);
    const bool c_emit = !GUARDED || t >= 5;
    // All tensor-memory reads of the step are requested HERE -- z rows t (stage A) and t-5 (stage C) and the 12 accumulator
    // columns of h2 row t-4, whose MMAs were triggered two steps ago -- and waited for at the end of stage B, which needs
    // none of them: ~110 instructions of conv-1 / conv-2 work cover the tcgen05.ld latency.
    zs.commit();
    uint32_t lz[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, ld[12] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (c_fma) {
        const int j = t - 4;
        mbar_wait(wk.dfull + (uint32_t)(j & 1) * 8u, (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        tmem_issue_ld12(wk.d_taddr + (uint32_t)(j & 1) * 16u, ld);
    }
    if (do_a) zs.issue_ld(t, lz[0], lz[1], lz[2], lz[3]);
    if (c_emit) zs.issue_ld(t - 5, lz[4], lz[5], lz[6], lz[7]);
    // ---------------- stage B
    Acc4 fin = b_old;
    if (b_fma) {
        const int i = t - 1;
        float2 xin[3];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) xin[dx] = wk.xr[i & 1][lane + dx];
        Acc4 nold, nmid;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            fin.v[o] = ffma2(xin[0], ld2(&P.w1[2][0][o][0]), b_old.v[o]);
            nold.v[o] = ffma2(xin[0], ld2(&P.w1[1][0][o][0]), b_mid.v[o]);
            nmid.v[o] = ffma2(xin[0], ld2(&P.w1[0][0][o][0]), zero2);
        }
#pragma unroll
        for (int dx = 1; dx < 3; ++dx) {
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                fin.v[o] = ffma2(xin[dx], ld2(&P.w1[2][dx][o][0]), fin.v[o]);
                nold.v[o] = ffma2(xin[dx], ld2(&P.w1[1][dx][o][0]), nold.v[o]);
                nmid.v[o] = ffma2(xin[dx], ld2(&P.w1[0][dx][o][0]), nmid.v[o]);
            }
        }
        b_old = nold;
        b_mid = nmid;
    } else if (GUARDED) {
        b_old = b_mid;   // t = 33: the zero row below the patch contributes nothing
    }
    if (b_emit) {
        const int r = t - 2;
        float h1[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) h1[o] = fmaxf(fin.v[o].x + fin.v[o].y + P.b1[o], 0.f);            // BN folded, ReLU
        const float2 h01 = make_float2(h1[0], h1[1]), h23 = make_float2(h1[2], h1[3]);
        float c2[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float2 u = ffma2(h01, w2c[o][0], zero2);
            u = ffma2(h23, w2c[o][1], u);
            c2[o] = u.x + u.y + P.b2[o];
        }
        uint32_t hi01, lo01, hi23, lo23;                       // ReLU is part of the split
        relu_split2(c2[0], c2[1], hi01, lo01);
        relu_split2(c2[2], c2[3], hi23, lo23);
        const int so = (r & 1) * A_STAGE;
        uint4 *ar = wk.a_row + so, *arr = wk.a_right + so, *arl = wk.a_left + so;
        const uint4 u = make_uint4(hi01, hi23, lo01, lo23);
        const uint2 h = make_uint2(hi01, hi23);
        ar[1 * 128] = u;                                               // centre tap of my own pixel
        reinterpret_cast<uint2*>(ar + 3 * 128)[1] = h;
        arr[0 * 128] = u;                                              // left tap of the pixel to my right
        reinterpret_cast<uint2*>(arr + 3 * 128)[0] = h;
        arl[2 * 128] = u;                                              // right tap of the pixel to my left
        reinterpret_cast<uint2*>(arl + 4 * 128)[0] = h;
    }
    // The loaded registers pass THROUGH the wait statement: no consumer can be scheduled above it.  It also precedes the
    // arrive below: the MMAs that arrive releases overwrite the accumulator stage this step has just read.
    tmem_wait_ld20(lz, ld);
    if (b_emit) {
        fence_async_smem();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
        tc_fence_before();       // orders the tcgen05.ld above before the MMAs of whichever thread issues them
        __syncwarp();
        // The four warps of a group take turns at issuing: the warp whose turn it is waits until all four have handed in
        // the row's A tile, then runs K = 48 as three M128 N16 K16 MMAs into the row's accumulator stage and commits them
        // to its "accumulators complete" barrier.  No issuer warp (a fifth warp on a scheduler would cap the workers at 102
        // registers), no polling.
        const int r = t - 2;
        const uint32_t st = (uint32_t)r & 1u;
        if (lane == 0) mbar_arrive(wk.afull + st * 8u);
        if ((r & 3) == wk.quarter) {
            mbar_wait(wk.afull + st * 8u, ((uint32_t)r >> 1) & 1u);
            tc_fence_after();
            const uint32_t a_addr = wk.a_tile + st * (uint32_t)(A_STAGE * 16), d = wk.d_mma + st * 16u;
            if (elect_one()) {
                mma_ss(d, make_desc(a_addr, 2048u, 128u), make_desc(b_addr, 256u, 128u), 0u);
                mma_ss(d, make_desc(a_addr + 4096u, 2048u, 128u), make_desc(b_addr + 512u, 256u, 128u), 1u);
                mma_ss(d, make_desc(a_addr + 8192u, 2048u, 128u), make_desc(b_addr + 1024u, 256u, 128u), 1u);
                mma_commit(wk.dfull + st * 8u);
            }
            __syncwarp();
        }
    }
    // ---------------- stage A
    if (do_a) {
        float4 z = make_float4(__uint_as_float(lz[0]), __uint_as_float(lz[1]), __uint_as_float(lz[2]), __uint_as_float(lz[3]));
        if (INV && has_mix) {
            z = mix4r(z, am);                                   // Conv2d1x1._inverse, layers.py:117-119
            zs.store(t, z);
        }
        wk.xr[t & 1][lane + 1] = make_float2(z.x, z.y);
    }
    // ---------------- stage C
    float h3[4];
    if (c_fma) {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            h3[o] = c_old[o] + __uint_as_float(ld[8 + o]);      // dy = 2: input row j = t-4 closes output row j - 1
            c_old[o] = c_mid[o] + __uint_as_float(ld[4 + o]);   // dy = 1 (carries the bias through the one-hot slots)
            c_mid[o] = __uint_as_float(ld[o]);                  // dy = 0: opens output row j + 1
        }
    } else {
#pragma unroll
        for (int o = 0; o < 4; ++o) h3[o] = c_old[o];   // t = 36: the zero row below the patch contributes nothing
    }
    if (c_emit) {
        const int q = t - 5;
        if (GUARDED && (q == 0 || q == 31)) {   // row class of the edge-indicator bias: the MMA applied the "middle" one
            const int rc = q == 0 ? 0 : 2, cc = lane == 0 ? 0 : (lane == 31 ? 2 : 1);
#pragma unroll
            for (int o = 0; o < 4; ++o) h3[o] += P.b3[rc][cc][o] - P.b3[1][cc][o];
        }
        // shift = h3[0:2], log_scale = scale * tanh(h3[2:4])                    (layers.py:362 / :342)
        const float ls0 = P.scale * fast_tanh(h3[2]);
        const float ls1 = P.scale * fast_tanh(h3[3]);
        float4 z = make_float4(__uint_as_float(lz[4]), __uint_as_float(lz[5]), __uint_as_float(lz[6]), __uint_as_float(lz[7]));
        if (INV) {
            z.z = fmaf(z.z, fast_exp(ls0), h3[0]);                               // layers.py:363-367
            z.w = fmaf(z.w, fast_exp(ls1), h3[1]);
            ldj += ls0 + ls1;                                                    // layers.py:372
        } else {
            z.z = (z.z - h3[0]) * fast_exp(-ls0);                                // layers.py:343-347
            z.w = (z.w - h3[1]) * fast_exp(-ls1);
            ldj -= ls0 + ls1;                                                    // layers.py:352
            if (has_mix) z = mix4r(z, am);                                       // Conv2d1x1._forward, layers.py:113-114
        }
        zs.store(q, z);
    }
    __syncwarp();
}

template <bool INV, class CP>
__device__ __forceinline__ void hyb_pass(const CP& P, const Worker& wk, const ZDual& zs, const uint32_t b_addr, const int lane, float& ldj) {
    const bool has_mix = P.has_mix != 0;
    const float rz = wk.xr[0][0].x;   // a 0.0f only known at run time: keeps the values below per-thread (see load_mix_regs)
    float2 am[4][2];
    load_mix_regs<INV>(P, am, rz);
    float2 w2c[4][2];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        w2c[o][0] = make_float2(P.w2[o][0] + rz, P.w2[o][1] + rz);
        w2c[o][1] = make_float2(P.w2[o][2] + rz, P.w2[o][3] + rz);
    }
    Acc4 b_old, b_mid;
    float c_old[4], c_mid[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        b_old.v[o] = b_mid.v[o] = make_float2(0.f, 0.f);
        c_old[o] = c_mid[o] = 0.f;
    }
#pragma unroll 1
    for (int t = 0; t < 37; ++t) {
        if (t >= 6 && t < 32) hyb_step<INV, false>(P, wk, zs, b_addr, lane, t, has_mix, b_old, b_mid, c_old, c_mid, ldj, am, w2c);
        else                  hyb_step<INV, true>(P, wk, zs, b_addr, lane, t, has_mix, b_old, b_mid, c_old, c_mid, ldj, am, w2c);
    }
}

#define NFH_FAST_SLOTS 8
template <bool INV>
__device__ __forceinline__ void hyb_dispatch(const NfModelParams& mp, const Worker& wk, const ZDual& zs, const uint32_t b0, int lane, float& ldj, int slot) {
    switch (slot) {
#define NFH_CASE(K) case K: hyb_pass<INV>(mp.cp[K], wk, zs, b0 + K * (uint32_t)(A_CHUNKS * 256), lane, ldj); break;
        NFH_CASE(0) NFH_CASE(1) NFH_CASE(2) NFH_CASE(3) NFH_CASE(4) NFH_CASE(5) NFH_CASE(6) NFH_CASE(7)
#undef NFH_CASE
        default: hyb_pass<INV>(mp.cp[slot], wk, zs, b0 + (uint32_t)slot * (uint32_t)(A_CHUNKS * 256), lane, ldj); break;   // run-time slot: per-thread LDC weight fetches
    }
}

template <bool INV>
__global__ void __launch_bounds__(THREADS, 1)
nf_chain_hyb_kernel(const __grid_constant__ NfModelParams mp, const NfChainArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler too

    // ---- CTA prologue: B tiles of every coupling, A tiles zeroed + their constant one-hot chunk, barriers, tensor memory
    {
        int n_cp = 0;
        for (int l = 0; l < mp.n_layers; ++l)
            if (mp.op[l] == NF_KOP_COUPLING && mp.slot[l] + 1 > n_cp) n_cp = mp.slot[l] + 1;
        uint16_t* b16 = reinterpret_cast<uint16_t*>(&S.b[0][0][0]);
        for (int e = tid; e < n_cp * (A_CHUNKS * 16 * 8); e += THREADS) {
            const int s = e / (A_CHUNKS * 128), rem = e - s * (A_CHUNKS * 128), j = rem >> 7, n = (rem >> 3) & 15, kk = rem & 7;
            b16[e] = half_bits(b_value(mp.cp[s], j, n, kk));   // [s][j][n][kk] is exactly the K-major no-swizzle layout
        }
        uint4* a4 = &S.a[0][0][0][0];
        for (int e = tid; e < GROUPS * 2 * A_STAGE; e += THREADS) {
            const int j = (e >> 7) % A_CHUNKS, col = e & 31;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (j == 5) {   // column class one-hot (fp16 1.0 = 0x3C00): [left, mid, right, 0] twice
                const uint32_t x = col == 0 ? 0x00003C00u : (col == 31 ? 0u : 0x3C000000u), y = col == 31 ? 0x00003C00u : 0u;
                v = make_uint4(x, y, x, y);
            }
            a4[e] = v;
        }
        for (int e = tid; e < WORKERS * 2 * 34; e += THREADS) (&S.xr[0][0][0])[e] = make_float2(0.f, 0.f);
        if (tid < GROUPS * 2) {
            mbar_init(smem_u32(&S.dfull[0][0] + tid), 1);
            mbar_init(smem_u32(&S.afull[0][0] + tid), 4);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    const int g = warp >> 2, q = warp & 3;
    const ZDual zs = {S.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((g < TMEM_GROUPS ? g : 0) * 128), &S.z[q][lane], g >= TMEM_GROUPS};
    Worker wk;
    wk.a_row = &S.a[g][0][0][q * 32 + lane];
    wk.a_right = lane < 31 ? wk.a_row + 1 : &S.a[GROUPS][0][0][2 * warp];       // every warp has its own two dump rows
    wk.a_left = lane > 0 ? wk.a_row - 1 : &S.a[GROUPS][0][0][2 * warp + 1];
    wk.xr = S.xr[warp];
    wk.a_tile = smem_u32(&S.a[g][0][0][0]);
    wk.afull = smem_u32(&S.afull[g][0]);
    wk.quarter = q;
    wk.dfull = smem_u32(&S.dfull[g][0]);
    wk.d_mma = S.tmem_base + D_COL0 + (uint32_t)g * 32u;
    wk.d_taddr = wk.d_mma + ((uint32_t)(q * 32) << 16);
    const uint32_t b0 = smem_u32(&S.b[0][0][0]);

    // CTA-uniform patch loop (see nf_chain_kernel): trailing warps without a patch recompute the last one
    const long long stride = (long long)gridDim.x * WORKERS;
    for (long long base = (long long)blockIdx.x * WORKERS; base < a.n; base += stride) {
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
        for (int l = l0; l != l1; l += dl) {
            const int op = mp.op[l], slot = mp.slot[l];
            switch (op) {
                case NF_KOP_COUPLING: hyb_dispatch<INV>(mp, wk, zs, b0, lane, ldj, slot); break;
                case NF_KOP_MIX: mix_pass<INV>(mp.mix[slot], zs); break;
                case NF_KOP_SDN:
                    sdn_pass<INV>(reinterpret_cast<const float4*>(a.y) + p * NF_PIXELS, mp.sc[slot].t[row][0], mp.sc[slot].t[row][1], zs, lane, ldj);
                    break;
                case NF_KOP_GAIN:
                    gain_pass<INV>(mp.sc[slot].t[row][0], mp.sc[slot].t[row][1], mp.sc[slot].t[row][2], zs, lane, ldj);
                    break;
                default: break;
            }
            __syncthreads();   // layer boundary: keeps the warps in the same loop body (instruction cache)
        }

        // ---- epilogue: store the patch, reduce log-det / prior / latent statistics (same fixed tree as nf_chain_kernel)
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
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(S.tmem_base), "r"(512u) : "memory");
}

}  // namespace hyb

bool hybrid_program_supported(const NfModelParams& mp, const NfChainArgs& a) {
    if (a.bn_stage != 0) return false;   // batch-statistics probes run on the all-fp32 kernel
    for (int l = a.first_layer; l < a.last_layer; ++l)
        if (mp.op[l] == NF_KOP_COUPLING) return true;
    return false;
}

template <bool INV>
static cudaError_t launch_hyb(const NfModelParams& mp, const NfChainArgs& args, int num_sms, cudaStream_t stream) {
    static bool attr_done[NF_MAX_DEVICES] = {};   // per device
    const int dev = device_slot();
    if (!attr_done[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(hyb::nf_chain_hyb_kernel<INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(hyb::Smem));
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    long long ctas = (args.n + hyb::WORKERS - 1) / hyb::WORKERS;
    if (ctas > num_sms) ctas = num_sms;
    hyb::nf_chain_hyb_kernel<INV><<<(unsigned)ctas, hyb::THREADS, sizeof(hyb::Smem), stream>>>(mp, args);
    return cudaGetLastError();
}

cudaError_t launch_chain_hybrid(const NfModelParams& mp, const NfChainArgs& args, bool inverse, int num_sms, cudaStream_t stream) {
    if (args.n <= 0) return cudaSuccess;
    return inverse ? launch_hyb<true>(mp, args, num_sms, stream) : launch_hyb<false>(mp, args, num_sms, stream);
}

}  // namespace nf
