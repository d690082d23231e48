// Wide coupling nets (width 32 / 64 / 128) on the 5th-generation tensor cores: real_nvp_conv_template(width)
// (layers.py:452-498; `--width` in sidd/ArgParser.py:43, "for Noise Flow it is 32": job_noise_flow.sh:19) with all three
// convolutions of every coupling as tcgen05.mma GEMMs, accumulators AND activations in tensor memory.
//
// The mapping.  A 128-pixel tile (4 image rows) is one M = 128 MMA tile: TMEM lane m <-> pixel m, for the accumulators
// D (fp32, one column per output channel) and -- because tcgen05.mma takes its A operand from tensor memory too -- for
// the activations: the thread that owns pixel m reads its accumulator row with tcgen05.ld, applies bias / ReLU, splits
// the result into bf16 (hi, lo) and writes it straight back into ITS OWN lane as the next GEMM's A operand
// (tcgen05.st; one 32-bit column = two consecutive K elements; validated by tools/tc_probe2.cu).  Activations never
// touch shared memory; weights (B operands, K-major no-swizzle [K/8][N][8] bf16) are staged once per coupling and CTA
// by the TMA engine (cp.async.bulk + mbarrier complete_tx) into a double-buffered block.
//
//   conv-1  3x3 SAME 2 -> W : A1 = im2col of the pixel's 3x3 x0 neighbourhood, K = 64:
//                             [x_hi 18 | x_lo 18 | x_hi 18 | 1 1 | 0..] x [W1_hi | W1_hi | W1_lo | b_hi b_lo]      4 MMAs
//   conv-2  1x1 W -> W      : A2 = [h1_hi W | h1_lo W];  per 16 channels  hi x [W2_hi | W2_lo] (N = 2W) + lo x W2_hi
//                             (W = 128: three N = W MMAs), bias through the constant-one slots of A1
//   conv-3  3x3 W -> 4      : as a 1x1 GEMM with N = 9 taps x 4 outputs (hi | lo = 96 columns), K = W, followed by a
//                             (column dy*16 + dx*4 + o) shifted sum on the CUDA cores: out(r, c) = sum_taps D3[(r + dy - 1, c + dx - 1), tap]
//                             (horizontal neighbours by warp shuffle, vertical by three ordered read-modify-write
//                             rounds on a 16 KB accumulator image) -- the hidden image never has to leave its lanes.
// fp32-grade accuracy from the bf16 (hi, lo) split of both operands: a w ~= a_hi w_hi + a_lo w_hi + a_hi w_lo, fp32
// accumulation in TMEM (dropped term 2^-16 relative per product).
//
// Work decomposition: G independent GROUPS per CTA (one CTA per SM), each owning one patch at a time (patch and
// accumulator image in shared memory, its own TMEM columns and MMA mbarrier); a group is 128 x H threads, the H warps
// that share a TMEM lane quarter split the channels.  One group's CUDA-core epilogue overlaps another group's MMAs
// (W = 32: G = 4, W = 64: G = 2, W = 128: G = 1, H = 4 / G).  A 17th warp is the TMA producer.
//
// Measured tcgen05 facts this is built on (profiles/r04_tc_probe2.log): an M = 128, K = 16 MMA costs max(46,
// N / 2) cycles whether A comes from shared or tensor memory -- so products are batched along N where that is free --
// and tcgen05.ld moves ~300 B/cycle/SM.
//
// Reference semantics as nf_wide.cu / nf_coupling.cuh (layers.py:117-130, 333-375, 452-498, 555-583, 651-674;
// noise_flow_model.py:394-480).
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


#ifndef NF_WTC_NB2
#define NF_WTC_NB2 1     // conv-2 (W <= 64): hi x [W2_hi | W2_lo] as one N = 2W MMA (D2 = [hh | hl], summed in the epilogue)
#endif
#ifndef NF_WTC_NB3
#define NF_WTC_NB3 1     // conv-3: hi x [W3_hi | W3_lo] as one N = 96 MMA (D3 = [hi part | lo part], summed in the epilogue)
#endif

template <int W>
struct Cfg {
    static_assert(W == 32 || W == 64 || W == 128, "resident-weights kernel: width 32 / 64 / 128");
    static constexpr int G = W == 32 ? 4 : (W == 64 ? 2 : 1);   // groups (patches in flight) per CTA
    static constexpr int H = 4 / G;                             // warps per TMEM lane quarter (channel split)
    static constexpr int GT = 128 * H;                          // threads per group
    static constexpr bool NB = W <= 64 && NF_WTC_NB2;           // conv-2 products batched along N (D2 = [hh | hl])
    static constexpr bool NB3 = NF_WTC_NB3 != 0;                // conv-3 products batched along N
    // TMEM columns of a group: A0 = A1 / A3, D = D1 / D2 (and D3 = 96 columns from D on), A2
    static constexpr int C_A0 = 0, C_D = W, C_A2 = C_D + (NB ? 2 * W : W), GROUP_COLS = C_A2 + W;
    static_assert(C_A2 + W - C_D >= 96, "D3 needs 96 columns behind D");
    static_assert(G * GROUP_COLS <= 512, "TMEM columns");
    static constexpr int NBUF = W <= 64 ? 2 : 1;                // weight blocks in flight
    using L = NfWideTcLayout;
    static constexpr int BLOCK_BYTES = NfWideTcLayout::block_bytes(W);
};


template <int W>
struct __align__(128) Smem {
    unsigned char wbuf[Cfg<W>::NBUF][Cfg<W>::BLOCK_BYTES];
    GroupSmem grp[Cfg<W>::G];
    uint64_t full[2], empty[2];
    uint32_t tmem_base;
    uint32_t pad_[3];
};


// ---------------------------------------------------------------------------------------------- one coupling
// STAGE 0: the coupling.  STAGE 1 / 2 (batch-statistics probes): accumulate per-channel sum and sum of squares of the
// conv-1 / conv-2 output before BatchNorm into Gs.sacc and leave z untouched (the host folds that BatchNorm as identity).
template <int W, bool INV, int STAGE>
__device__ __forceinline__ void tc_coupling(const unsigned char* wb, GroupSmem& Gs, const int g, const int i, const uint32_t tmem_g,
                                            uint32_t& mphase, float& ldj) {
    using C = Cfg<W>;
    using L = NfWideTcLayout;
    constexpr int GT = C::GT, H = C::H;
    const int lane = i & 31, wq = (i >> 5) & 3, h = i >> 7;
    // the group's copy of the fp32 header: the weight block may be released before the affine update is done
    if (i < 128) Gs.hdr[i] = reinterpret_cast<const float*>(wb)[i];
    group_barrier(g, GT);
    const float* hdr = Gs.hdr;
    const bool has_mix = hdr[L::H_META] != 0.f;
    // (inverse) 1x1 mix; conv-3 accumulator image starts from the folded bias by (row class, column class)
#pragma unroll 2
    for (int k = 0; k < NF_PIXELS / GT; ++k) {
        const int px = k * GT + i, r = px >> 5, c = px & 31;
        if (INV && has_mix && STAGE == 0) Gs.z[px] = mix4(Gs.z[px], hdr + L::H_A);
        const int rc = r == 0 ? 0 : (r == 31 ? 2 : 1), cc = c == 0 ? 0 : (c == 31 ? 2 : 1);
        Gs.pre[px] = *reinterpret_cast<const float4*>(hdr + L::H_B3 + (rc * 3 + cc) * 4);
    }
    group_barrier(g, GT);
    float acc_s = 0.f, acc_q = 0.f;    // probes
    const bool mix_on_the_fly = STAGE != 0 && INV && has_mix;   // a probe must not modify z: mix while building A1

    const uint32_t wb_addr = smem_u32(wb);
    const uint32_t mbar = smem_u32(&Gs.mbar);
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
    const uint32_t tA0 = tmem_g + C::C_A0, tD = tmem_g + C::C_D, tA2 = tmem_g + C::C_A2;

#pragma unroll 1
    for (int t = 0; t < 8; ++t) {
        const int r = 4 * t + wq;
        // ---------------- A1: im2col of x0 (hi | lo | hi | 1 1 | 0), 32 columns, by the h == 0 thread of the pixel
        if (h == 0) {
            uint32_t a1[32];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int rr = r + dy - 1, cc = lane + dx - 1;
                    float2 x0 = make_float2(0.f, 0.f);
                    if (rr >= 0 && rr <= 31 && cc >= 0 && cc <= 31) {
                        if (mix_on_the_fly) { const float4 zz = mix4(Gs.z[rr * 32 + cc], hdr + L::H_A); x0 = make_float2(zz.x, zz.y); }
                        else x0 = *reinterpret_cast<const float2*>(&Gs.z[rr * 32 + cc]);
                    }
                    uint32_t hi, lo;
                    split2(x0.x, x0.y, hi, lo);
                    a1[dy * 3 + dx] = hi;
                    a1[9 + dy * 3 + dx] = lo;
                    a1[18 + dy * 3 + dx] = hi;
                }
            a1[27] = 0x3F803F80u;   // (1.0, 1.0): bias slots of conv-1 and conv-2
            a1[28] = a1[29] = a1[30] = a1[31] = 0u;
            uint32_t lo16[16], hi16[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) { lo16[k] = a1[k]; hi16[k] = a1[16 + k]; }
            tmem_st16(tA0 + lane_sel, lo16);
            tmem_st16(tA0 + lane_sel + 16u, hi16);
            tmem_wait_st();
        }
        tc_fence_before();
        group_barrier(g, GT);
        // ---------------- conv-1: D1[128 x W] = A1[128 x 64] . B1
        if (i == 0) {
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 4; ++j)
                mma_ts(tD, tA0 + 8u * j, make_desc(wb_addr + L::off_b1() + (uint32_t)j * 2u * W * 16u, W * 16u, 128u), idesc(W), j > 0 ? 1u : 0u);
            mma_commit(mbar);
        }
        group_wait_mma(mbar, mphase, g, i, GT);
        tc_fence_after();
        if (STAGE == 1) { probe_accumulate<false>(tD + lane_sel + 32u * h, 0u, lane, acc_s, acc_q); continue; }
        // ---------------- epilogue 1: ReLU, split, A2 (this thread: channels [32 h, 32 h + 32))
        relu_split_store<false>(tD + lane_sel + 32u * h, 0u, tA2 + lane_sel + 16u * h, tA2 + lane_sel + W / 2 + 16u * h);
        tc_fence_before();
        group_barrier(g, GT);
        // ---------------- conv-2: D2 = A2 . B2 (+ bias through the constant-one slots of A1, K chunk 3)
        if (i == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < W / 16; ++s) {
                const uint32_t bk = wb_addr + L::off_b2(W) + (uint32_t)s * 2u * (2u * W * 16u);
                if (C::NB) {
                    mma_ts(tD, tA2 + 8u * s, make_desc(bk, 2u * W * 16u, 128u), idesc(2 * W), s > 0 ? 1u : 0u);          // hi x [hi | lo]
                    mma_ts(tD, tA2 + W / 2 + 8u * s, make_desc(bk, 2u * W * 16u, 128u), idesc(W), 1u);                    // lo x hi
                } else {
                    mma_ts(tD, tA2 + 8u * s, make_desc(bk, 2u * W * 16u, 128u), idesc(W), s > 0 ? 1u : 0u);              // hi x hi
                    mma_ts(tD, tA2 + W / 2 + 8u * s, make_desc(bk, 2u * W * 16u, 128u), idesc(W), 1u);                    // lo x hi
                    mma_ts(tD, tA2 + 8u * s, make_desc(bk + W * 16u, 2u * W * 16u, 128u), idesc(W), 1u);                  // hi x lo
                }
            }
            mma_ts(tD, tA0 + 24u, make_desc(wb_addr + L::off_bb2(W), W * 16u, 128u), idesc(W), 1u);
            mma_commit(mbar);
        }
        group_wait_mma(mbar, mphase, g, i, GT);
        tc_fence_after();
        if (STAGE == 2) { probe_accumulate<C::NB>(tD + lane_sel + 32u * h, tD + lane_sel + W + 32u * h, lane, acc_s, acc_q); continue; }
        // ---------------- epilogue 2: ReLU, split, A3 (over A1)
        relu_split_store<C::NB>(tD + lane_sel + 32u * h, tD + lane_sel + W + 32u * h, tA0 + lane_sel + 16u * h,
                                tA0 + lane_sel + W / 2 + 16u * h);
        tc_fence_before();
        group_barrier(g, GT);
        // ---------------- conv-3 as a 1x1 GEMM: D3[128 x (36 hi-part | 36 lo-part)] = A3 . [B3_hi | B3_lo]
        if (i == 0) {
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < W / 16; ++s) {
                const uint32_t bk = wb_addr + L::off_b3(W) + (uint32_t)s * 2u * (96u * 16u);
                if (C::NB3) {
                    mma_ts(tD, tA0 + 8u * s, make_desc(bk, 96u * 16u, 128u), idesc(96), s > 0 ? 1u : 0u);                 // hi x [hi | lo]
                    mma_ts(tD, tA0 + W / 2 + 8u * s, make_desc(bk, 96u * 16u, 128u), idesc(48), 1u);                       // lo x hi
                } else {
                    mma_ts(tD, tA0 + 8u * s, make_desc(bk, 96u * 16u, 128u), idesc(48), s > 0 ? 1u : 0u);                 // hi x hi
                    mma_ts(tD, tA0 + W / 2 + 8u * s, make_desc(bk, 96u * 16u, 128u), idesc(48), 1u);                       // lo x hi
                    mma_ts(tD, tA0 + 8u * s, make_desc(bk + 48u * 16u, 96u * 16u, 128u), idesc(48), 1u);                   // hi x lo
                }
            }
            mma_commit(mbar);
        }
        group_wait_mma(mbar, mphase, g, i, GT);
        tc_fence_after();
        // ---------------- epilogue 3: shifted sum.  Pixel (r, c), tap (dy, dx) contributes to output (r - dy + 1, c - dx + 1):
        // horizontal neighbours by warp shuffle (s_dy = the three dx taps of row r gathered at their output column), vertical
        // ones through a 4-row exchange buffer -- ONE barrier; rows of the neighbouring tiles are updated in `pre` directly.
        float s_dy[3][4];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            if (h == (dy * H) / 3) {
                uint32_t a[16], b[16];
                tmem_ld16(tD + lane_sel + 16u * dy, a);
                if (C::NB3) tmem_ld16(tD + lane_sel + 48u + 16u * dy, b);
                tmem_wait_ld();
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    float d0 = __uint_as_float(a[o]), d1 = __uint_as_float(a[4 + o]), d2 = __uint_as_float(a[8 + o]);
                    if (C::NB3) { d0 += __uint_as_float(b[o]); d1 += __uint_as_float(b[4 + o]); d2 += __uint_as_float(b[8 + o]); }
                    float fl = __shfl_up_sync(0xffffffffu, d0, 1);      // dx = 0 of column c - 1 lands here
                    float fr = __shfl_down_sync(0xffffffffu, d2, 1);    // dx = 2 of column c + 1 lands here
                    if (lane == 0) fl = 0.f;
                    if (lane == 31) fr = 0.f;
                    s_dy[dy][o] = d1 + (fl + fr);
                }
                if (dy == 0) Gs.ex[0][wq][lane] = make_float4(s_dy[0][0], s_dy[0][1], s_dy[0][2], s_dy[0][3]);   // goes to row r + 1
                if (dy == 2) Gs.ex[1][wq][lane] = make_float4(s_dy[2][0], s_dy[2][1], s_dy[2][2], s_dy[2][3]);   // goes to row r - 1
            }
        }
        tc_fence_before();   // the next tile's MMAs / stores reuse these TMEM columns
        group_barrier(g, GT);
        if (h == (1 * H) / 3) {            // own row: dy = 1 of this row + dy = 0 of the row above + dy = 2 of the row below
            float4 acc = make_float4(s_dy[1][0], s_dy[1][1], s_dy[1][2], s_dy[1][3]);
            if (wq > 0) { const float4 u = Gs.ex[0][wq - 1][lane]; acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w; }
            if (wq < 3) { const float4 d = Gs.ex[1][wq + 1][lane]; acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w; }
            float4 p = Gs.pre[r * 32 + lane];
            p.x += acc.x; p.y += acc.y; p.z += acc.z; p.w += acc.w;
            Gs.pre[r * 32 + lane] = p;
        }
        if (h == 0 && wq == 3 && r < 31) {            // dy = 0 of the tile's last row -> first row of the next tile
            float4 p = Gs.pre[(r + 1) * 32 + lane];
            p.x += s_dy[0][0]; p.y += s_dy[0][1]; p.z += s_dy[0][2]; p.w += s_dy[0][3];
            Gs.pre[(r + 1) * 32 + lane] = p;
        }
        if (h == (2 * H) / 3 && wq == 0 && r > 0) {   // dy = 2 of the tile's first row -> last row of the previous tile
            float4 p = Gs.pre[(r - 1) * 32 + lane];
            p.x += s_dy[2][0]; p.y += s_dy[2][1]; p.z += s_dy[2][2]; p.w += s_dy[2][3];
            Gs.pre[(r - 1) * 32 + lane] = p;
        }
    }
    group_barrier(g, GT);   // `pre` is complete
    if (STAGE != 0) {   // lane l of warp (wq, h) owns channel 32 h + l of its image rows; the four row-warps meet in shared memory
        atomicAdd(&Gs.sacc[32 * h + lane], acc_s);
        atomicAdd(&Gs.sacc[W + 32 * h + lane], acc_q);
        return;
    }
    // ---------------- affine coupling update of z, log-det (layers.py:333-375)
    const float scale = hdr[L::H_META + 1];
#pragma unroll 2
    for (int k = 0; k < NF_PIXELS / GT; ++k) {
        const int px = k * GT + i;
        const float4 p = Gs.pre[px];
        float4 z = Gs.z[px];
        const float ls0 = scale * t_tanh(p.z), ls1 = scale * t_tanh(p.w);
        if (INV) {
            z.z = fmaf(z.z, t_exp(ls0), p.x);
            z.w = fmaf(z.w, t_exp(ls1), p.y);
            ldj += ls0 + ls1;
        } else {
            z.z = (z.z - p.x) * t_exp(-ls0);
            z.w = (z.w - p.y) * t_exp(-ls1);
            ldj -= ls0 + ls1;
            if (has_mix) z = mix4(z, hdr + L::H_AINV);
        }
        Gs.z[px] = z;
    }
}

template <int W, bool INV>
__global__ void __launch_bounds__(THREADS, 1)
nf_wide_tc_kernel(const __grid_constant__ CUtensorMap tmap, const NfWideProgram prog, const float* __restrict__ blob, const NfChainArgs a) {
    using C = Cfg<W>;
    constexpr int G = C::G, GT = C::GT, NBUF = C::NBUF;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<W>& S = *reinterpret_cast<Smem<W>*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool producer = tid >= COMPUTE_THREADS;
    const int g = producer ? 0 : tid / GT, i = tid % GT;

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&S.full[b]), 1u); mbar_init(smem_u32(&S.empty[b]), (uint32_t)G); }
        for (int k = 0; k < G; ++k) mbar_init(smem_u32(&S.grp[k].mbar), 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const long long per_round = (long long)gridDim.x * G;
    const long long rounds = (a.n + per_round - 1) / per_round;
    const int n_l = a.last_layer - a.first_layer;

    if (producer) {
        // ---- TMA producer: one weight block per coupling execution, in the order every group consumes them
        uint32_t item = 0;       // the whole warp walks the schedule (it meets the CTA barrier converged); lane 0 issues
        for (long long rd = 0; rd < rounds; ++rd)
            for (int step = 0; step < n_l; ++step) {
                const int l = INV ? a.first_layer + step : a.last_layer - 1 - step;
                if (prog.op[l] != NF_KOP_COUPLING) continue;
                const uint32_t b = item % NBUF, use = item / NBUF;
                if (lane == 0) {
                    if (use > 0) mbar_wait(smem_u32(&S.empty[b]), (use - 1u) & 1u);
                    const uint32_t fb = smem_u32(&S.full[b]);
                    mbar_arrive_expect_tx(fb, (uint32_t)C::BLOCK_BYTES);
                    // TMA tensor copies: the block is BLOCK_ROWS rows of the blob's 2-D view, fetched in boxes of BOX_ROWS
                    constexpr int BLOCK_ROWS = C::BLOCK_BYTES / (NF_TMA_ROW_FLOATS * 4), BOX_ROWS = W == 32 ? 62 : (W == 64 ? 77 : 217);
                    static_assert(BLOCK_ROWS % BOX_ROWS == 0, "box rows");
                    const int row0 = prog.off[l] / NF_TMA_ROW_FLOATS;
                    for (int r = 0; r < BLOCK_ROWS; r += BOX_ROWS)
                        tma_load_rows(smem_u32(&S.wbuf[b][(size_t)r * NF_TMA_ROW_FLOATS * 4]), &tmap, row0 + r, fb);
                }
                __syncwarp();
                ++item;
            }
    } else {
        GroupSmem& Gs = S.grp[g];
        const uint32_t tmem_g = S.tmem_base + (uint32_t)(g * C::GROUP_COLS);
        uint32_t item = 0, mphase = 0;
        if (i < 256) Gs.sacc[i] = 0.f;
        for (long long rd = 0; rd < rounds; ++rd) {
            const long long p = (rd * gridDim.x + blockIdx.x) * G + g;
            const bool valid = p < a.n;
            int row = 0;
            if (valid) {
                row = a.rows ? a.rows[p] : a.default_row;
                row = min(max(row, 0), NF_MAX_ROWS - 1);
#pragma unroll 2
                for (int k = 0; k < NF_PIXELS / GT; ++k) {
                    const int px = k * GT + i;
                    float4 v;
                    if (a.in) {
                        v = __ldcs(reinterpret_cast<const float4*>(a.in) + p * NF_PIXELS + px);
                        if (!INV) { v.x *= a.temp; v.y *= a.temp; v.z *= a.temp; v.w *= a.temp; }   // noise_flow_model.py:501
                    } else {
                        v = philox_normal4(a.seed, a.offset, a.patch_base + (unsigned long long)p, (unsigned int)px);
                        v.x *= a.temp; v.y *= a.temp; v.z *= a.temp; v.w *= a.temp;
                    }
                    Gs.z[px] = v;
                }
            }
            group_barrier(g, GT);
            float ldj = 0.f;
            for (int step = 0; step < n_l; ++step) {
                const int l = INV ? a.first_layer + step : a.last_layer - 1 - step;
                const int op = prog.op[l];
                const float* pb = blob + prog.off[l];
                if (op == NF_KOP_COUPLING) {
                    const uint32_t b = item % NBUF, use = item / NBUF;
                    mbar_wait(smem_u32(&S.full[b]), use & 1u);
                    const bool probe = a.bn_stage != 0 && step == n_l - 1;       // the last op executed is the coupling under probe
                    if (valid) {
                        if (!probe) tc_coupling<W, INV, 0>(S.wbuf[b], Gs, g, i, tmem_g, mphase, ldj);
                        else if (a.bn_stage == 1) tc_coupling<W, INV, 1>(S.wbuf[b], Gs, g, i, tmem_g, mphase, ldj);
                        else tc_coupling<W, INV, 2>(S.wbuf[b], Gs, g, i, tmem_g, mphase, ldj);
                    }
                    group_barrier(g, GT);          // every MMA that reads the block has completed, the header is copied
                    if (i == 0) mbar_arrive(smem_u32(&S.empty[b]));
                    ++item;
                    continue;
                }
                if (valid) {
                    if (op == NF_KOP_MIX) {
                        for (int k = 0; k < NF_PIXELS / GT; ++k) Gs.z[k * GT + i] = mix4(Gs.z[k * GT + i], pb + (INV ? 0 : 16));
                    } else if (op == NF_KOP_SDN) {
                        const float sa = __ldg(pb + row * 4), sb = __ldg(pb + row * 4 + 1);
                        const float4* yp = reinterpret_cast<const float4*>(a.y) + p * NF_PIXELS;
                        float acc = 0.f;
                        for (int k = 0; k < NF_PIXELS / GT; ++k) {
                            const float4 y = __ldg(yp + k * GT + i);
                            float4 z = Gs.z[k * GT + i];
                            const float v0 = fmaf(sa, y.x, sb), v1 = fmaf(sa, y.y, sb), v2 = fmaf(sa, y.z, sb), v3 = fmaf(sa, y.w, sb);
                            const float r0 = rsqrtf(v0), r1 = rsqrtf(v1), r2 = rsqrtf(v2), r3 = rsqrtf(v3);
                            if (INV) { z.x *= r0; z.y *= r1; z.z *= r2; z.w *= r3; }                       // SdnEx5.py:125-126
                            else     { z.x *= v0 * r0; z.y *= v1 * r1; z.z *= v2 * r2; z.w *= v3 * r3; }   // SdnEx5.py:106-107
                            acc += (__logf(v0) + __logf(v1)) + (__logf(v2) + __logf(v3));
                            Gs.z[k * GT + i] = z;
                        }
                        ldj += INV ? -0.5f * acc : 0.5f * acc;
                    } else if (op == NF_KOP_GAIN) {
                        const float mlt = INV ? __ldg(pb + row * 4 + 1) : __ldg(pb + row * 4);
                        for (int k = 0; k < NF_PIXELS / GT; ++k) {
                            float4 z = Gs.z[k * GT + i];
                            z.x *= mlt; z.y *= mlt; z.z *= mlt; z.w *= mlt;
                            Gs.z[k * GT + i] = z;
                        }
                        if (i == 0) ldj += INV ? __ldg(pb + row * 4 + 2) : -__ldg(pb + row * 4 + 2);
                    }
                }
                group_barrier(g, GT);
            }
            if (a.bn_stage != 0) {   // probe launch: publish this patch's per-channel sums, nothing else
                if (valid && i < 2 * W) { atomicAdd(a.bn_stats + i, (double)Gs.sacc[i]); Gs.sacc[i] = 0.f; }
                group_barrier(g, GT);
                continue;
            }
            // ---- epilogue: store the patch, reduce log-det / prior / latent statistics (fixed order)
            if (valid) {
                float s1 = 0.f, s2 = 0.f;
                float4* dst = a.out ? reinterpret_cast<float4*>(a.out) + p * NF_PIXELS : nullptr;
                for (int k = 0; k < NF_PIXELS / GT; ++k) {
                    const float4 z = Gs.z[k * GT + i];
                    if (dst) __stcs(dst + k * GT + i, z);
                    s1 += (z.x + z.y) + (z.z + z.w);
                    s2 = fmaf(z.x, z.x, fmaf(z.y, z.y, fmaf(z.z, z.z, fmaf(z.w, z.w, s2))));
                }
                ldj = wsum(ldj); s1 = wsum(s1); s2 = wsum(s2);
                if (lane == 0) { Gs.red[(i >> 5) * 4] = ldj; Gs.red[(i >> 5) * 4 + 1] = s1; Gs.red[(i >> 5) * 4 + 2] = s2; }
            }
            group_barrier(g, GT);
            if (valid && i == 0) {
                float t_ldj = 0.f, t1 = 0.f, t2 = 0.f;
                for (int k = 0; k < GT / 32; ++k) { t_ldj += Gs.red[k * 4]; t1 += Gs.red[k * 4 + 1]; t2 += Gs.red[k * 4 + 2]; }
                const float logdet = t_ldj + (INV ? a.ldj_const : -a.ldj_const) + (a.logdet_in ? a.logdet_in[p] : 0.f);
                if (a.logdet) a.logdet[p] = logdet;
                if (a.nll) a.nll[p] = -(logdet - 0.5f * (NF_DIMS * 1.8378770664093453f + t2));   // noise_flow_model.py:474-475,537-539
                if (a.sdz) {                                                                      // noise_flow_model.py:477-478
                    const float mean = t1 * (1.f / NF_DIMS);
                    a.sdz[p] = sqrtf(fmaxf(t2 * (1.f / NF_DIMS) - mean * mean, 0.f));
                }
            }
            group_barrier(g, GT);
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(S.tmem_base), "r"(512u) : "memory");
}

template <int W>
static cudaError_t launch_w(const NfWideProgram& prog, const float* blob, const NfChainArgs& a, bool inverse, int num_sms, cudaStream_t stream) {
    const size_t smem = sizeof(Smem<W>);
    cudaError_t e = cudaFuncSetAttribute(nf_wide_tc_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device: set on every launch
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(nf_wide_tc_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    CUtensorMap tmap;
    e = make_blob_tensor_map(blob, prog.blob_floats, wide_tc_box_rows(W), &tmap);
    if (e != cudaSuccess) return e;
    long long grid = (a.n + Cfg<W>::G - 1) / Cfg<W>::G;
    if (grid > num_sms) grid = num_sms;
    if (inverse) nf_wide_tc_kernel<W, true><<<(unsigned)grid, THREADS, smem, stream>>>(tmap, prog, blob, a);
    else nf_wide_tc_kernel<W, false><<<(unsigned)grid, THREADS, smem, stream>>>(tmap, prog, blob, a);
    return cudaGetLastError();
}

}  // namespace wtc

bool wide_tc_width_supported(int width) { return width == 32 || width == 64 || width == 128 || width == 256 || width == 512; }

cudaError_t launch_chain_wide_tc(const NfWideProgram& prog, const float* blob, const NfChainArgs& a, bool inverse, int num_sms,
                                 cudaStream_t stream) {
    if (a.n <= 0) return cudaSuccess;
    switch (prog.width) {
        case 32: return wtc::launch_w<32>(prog, blob, a, inverse, num_sms, stream);
        case 64: return wtc::launch_w<64>(prog, blob, a, inverse, num_sms, stream);
        case 128: return wtc::launch_w<128>(prog, blob, a, inverse, num_sms, stream);
        case 256:
        case 512: return launch_chain_wide_tcs(prog, blob, a, inverse, num_sms, stream);   // streamed weights (nf_wide_tcs.cu)
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace nf
