// Device-resident train step of the Noise Flow chain (C-ABI: nf_trainer_* in include/noiseflow_b200.h).
//
// What sess.run([train_op, loss, sd_z], is_training=True) does in the reference (train_noise_flow.py:50-77,
// 187-198) runs here without a single host round trip: every TF variable, the Adam slots and the step counter
// stay in device memory; the parameterisations that the host-synchronous path (nf_train.cu + train.py) handles
// on the CPU -- LU assembly of the 1x1 matrices, per-(camera, ISO) scale tables, BatchNorm batch statistics,
// the chain rules back to the LU / scale variables, Adam, the BatchNorm moving averages -- are small kernels
// next to the heavy passes, so a step has no cudaStreamSynchronize in it: one cooperative kernel (td_step_kernel) plus the
// reduce / chain-rule / Adam kernels when the batch is co-resident, ~60 per-pass launches otherwise.
//
// Mapping: ONE CTA OWNS ONE PATCH (8 warps, 16 selectable; warp w owns image rows w, w+NW, ...; lane = column).  A train batch
// is 138-207 patches per GPU (job_noise_flow.sh:37), far fewer than the 148 x 16 resident warps of the
// inference kernel, so the patch is split over a whole CTA to cut the latency of every pass by ~8x; images live
// in that CTA's shared memory; parameter-gradient partial sums go warp shuffle -> shared fp32 -> one fp64
// atomic per CTA and parameter.
//
// Math (identical to nf_train.cu, which stays as the independently verified reference implementation):
//   coupling, inverse direction, raw parameters:
//     z' = z_in.A ; x0 = z'[:2], x1 = z'[2:]
//     c1 = conv3x3_SAME(x0; W1) + b1 ; h1 = relu((c1-m1)/s1) ; c2 = h1.W2 + b2 ; h2 = relu((c2-m2)/s2)
//     h3 = (conv3x3_VALID(pad(h2) (+) ring; W3) + b3) * exp(3 logs) ; shift = h3[:2], raw = h3[2:]
//     ls = scale * tanh(raw) ; out = [x0, x1*exp(ls) + shift] ; ldj = sum ls
//   forward passes F1 (sums of c1), F2 (sums of c2), F3 (apply) ; backward passes B1, B2, B3 as in nf_train.cu.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <new>
#include <vector>

#include "../../include/noiseflow_b200.h"
#include "nf_kernels.h"
#include "nf_params.h"
#include "nf_train.h"
#include "nf_train_common.cuh"
#include "nf_rng.cuh"

namespace nf {

#define TD_THREADS 256
#define TD_WARPS (TD_THREADS / 32)
#define TD_MAX_OPS 40
#define TD_BN_MOMENTUM 0.1f     // layers.py:394-395

struct __align__(16) TdSmem {
    NfTrainCoupling P;                     // raw parameters + BatchNorm statistics in force
    float acc[16][NF_G_COUPLING_DOUBLES];  // per-warp partial sums, NF_G_* layout (a warp owns its row: no shared atomics)
    float red[16];                         // per-warp scratch of the log-det sum
    float4 zp[NF_PIXELS];                  // z' = z_in . A
    float4 c1[NF_PIXELS];                  // conv-1 output of the resident patch (kept across the passes of the fused kernel)
    float4 h2[34 * 34];                    // padded h2 image (ring = 0)
    float4 g[34 * 34];                     // padded gradient image
};

// K partial sums per lane -> lane l (l < K) ends up with the warp-wide total of v[l].  Transposed butterfly: at offset
// `half` a lane keeps the half of its values whose index has that bit equal to its own lane bit and trades the other
// half, so the value count halves with every step: P - 1 shuffles for P = 2^ceil(log2 K) values (offsets >= P are plain
// butterflies over the K values) instead of 5 K -- the parameter-gradient reductions were ~1000 SHFL per warp and patch
// in pass B1, half of its stall samples (profiles/r02_td_b1_train_207_ncu_full.txt).
template <int K>
__device__ __forceinline__ float warp_reduce_to_lanes(const float (&v)[K], int lane) {
    constexpr int P = K <= 1 ? 1 : K <= 2 ? 2 : K <= 4 ? 4 : K <= 8 ? 8 : K <= 16 ? 16 : 32;
    static_assert(K >= 1 && K <= 32, "one value per lane at most");
    float x[P];
#pragma unroll
    for (int i = 0; i < P; ++i) x[i] = i < K ? v[i] : 0.f;
#pragma unroll
    for (int half = 16; half >= P; half >>= 1)
#pragma unroll
        for (int i = 0; i < K; ++i) x[i] += __shfl_xor_sync(0xffffffffu, x[i], half);
#pragma unroll
    for (int half = P / 2; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float keep = up ? x[i + half] : x[i];
            const float send = up ? x[i] : x[i + half];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return x[0];      // lane l: total of v[l % P]
}
// accumulation of K consecutive slots into THIS WARP'S row of partial sums (lane l -> slot l; shared-memory float atomics are
// CAS loops on this architecture and were 6 % of the fused kernel's stall samples); td_flush adds the rows in a fixed order
template <int K>
__device__ __forceinline__ void cta_acc_vec(float* acc_row, const float (&v)[K], int lane) {
    const float tot = warp_reduce_to_lanes<K>(v, lane);
    if (lane < K) acc_row[lane] += tot;
}
__device__ __forceinline__ bool on_ring(int k) {
    const int R = k / 34, C = k - R * 34;
    return R == 0 || R == 33 || C == 0 || C == 33;
}

// conv-1 (3x3 SAME, 2 -> 4 channels) of the rows this warp owns, register-blocked over those rows: the 8 weights of a tap
// are read from shared memory once and applied to every row (9 x (2 + R) LDS.128 for R rows instead of 27 R), no divergent
// edge handling (column clamped, contribution masked).  The result is parked in S.c1; every later use in this pass -- and,
// in the fused kernel, in the following passes of the same coupling -- reads it back (each thread reads only what it wrote).
// conv-1 recomputed per pixel by all six passes was 13 % of the fused kernel's stall samples.
template <int NW>
__device__ __forceinline__ void td_conv1_rows(TdSmem& S, int warp, int lane) {
    constexpr int R = 32 / NW;
    float acc[R][4];
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[k][o] = S.P.b1[o];
#pragma unroll 1
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const float4 w0 = *reinterpret_cast<const float4*>(&S.P.w1[dy][dx][0][0]);
            const float4 w1 = *reinterpret_cast<const float4*>(&S.P.w1[dy][dx][1][0]);
            const int cc = lane + dx - 1, ccs = min(max(cc, 0), 31);
            const float cm = cc == ccs ? 1.f : 0.f;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int rr = warp + k * NW + dy - 1;
                if (rr < 0 || rr > 31) continue;                      // warp-uniform
                const float4 z = S.zp[rr * 32 + ccs];
                const float zx = z.x * cm, zy = z.y * cm;
                acc[k][0] = fmaf(zx, w0.x, fmaf(zy, w1.x, acc[k][0]));
                acc[k][1] = fmaf(zx, w0.y, fmaf(zy, w1.y, acc[k][1]));
                acc[k][2] = fmaf(zx, w0.z, fmaf(zy, w1.z, acc[k][2]));
                acc[k][3] = fmaf(zx, w0.w, fmaf(zy, w1.w, acc[k][3]));
            }
        }
#pragma unroll
    for (int k = 0; k < R; ++k) S.c1[(warp + k * NW) * 32 + lane] = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
}
__device__ __forceinline__ void td_c1(const TdSmem& S, int r, int lane, float (&c1)[4]) {
    const float4 v = S.c1[r * 32 + lane];
    c1[0] = v.x; c1[1] = v.y; c1[2] = v.z; c1[3] = v.w;
}
// h1 (post BatchNorm-1 + ReLU) and the normalised c2hat at pixel (r, lane) from the parked conv-1 output
__device__ __forceinline__ void td_net_to_c2hat(const TdSmem& S, int r, int lane, float (&c1hat)[4], float (&h1)[4], float (&c2hat)[4]) {
    float c1[4];
    td_c1(S, r, lane, c1);
#pragma unroll
    for (int o = 0; o < 4; ++o) { c1hat[o] = (c1[o] - S.P.m1[o]) * S.P.is1[o]; h1[o] = fmaxf(c1hat[o], 0.f); }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float c2 = S.P.b2[o];
#pragma unroll
        for (int i = 0; i < 4; ++i) c2 = fmaf(h1[i], S.P.w2[i][o], c2);
        c2hat[o] = (c2 - S.P.m2[o]) * S.P.is2[o];
    }
}

// conv-3 (edge-padded: 3x3 VALID over the zero-padded h2 image + the ring-indicator channel, 5 -> 4, layers.py:555-583,
// 651-674) of the rows this warp owns, register-blocked like conv-1: 9 x (5 + R) LDS.128 instead of 54 R.
template <int NW>
__device__ __forceinline__ void td_conv3_rows(const TdSmem& S, int warp, int lane, float (&pre)[32 / NW][4]) {
    constexpr int R = 32 / NW;
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
        for (int o = 0; o < 4; ++o) pre[k][o] = S.P.b3[o];
#pragma unroll 1
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            float4 w[5];
#pragma unroll
            for (int ci = 0; ci < 5; ++ci) w[ci] = *reinterpret_cast<const float4*>(&S.P.w3[dy][dx][ci][0]);
            const int C = lane + dx;
            const bool cedge = C == 0 || C == 33;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int Rr = warp + k * NW + dy;
                const float4 h = S.h2[Rr * 34 + C];
                const float ring = (cedge || Rr == 0 || Rr == 33) ? 1.f : 0.f;
                pre[k][0] += h.x * w[0].x + h.y * w[1].x + h.z * w[2].x + h.w * w[3].x + ring * w[4].x;
                pre[k][1] += h.x * w[0].y + h.y * w[1].y + h.z * w[2].y + h.w * w[3].y + ring * w[4].y;
                pre[k][2] += h.x * w[0].z + h.y * w[1].z + h.z * w[2].z + h.w * w[3].z + ring * w[4].z;
                pre[k][3] += h.x * w[0].w + h.y * w[1].w + h.z * w[2].w + h.w * w[3].w + ring * w[4].w;
            }
        }
}
// transposed conv-3: g_h2(r, c)[ci] = sum_{dy,dx,o} W3[dy][dx][ci][o] * g_pre3(r-dy+1, c-dx+1)[o] over the padded gradient image
template <int NW>
__device__ __forceinline__ void td_conv3t_rows(const TdSmem& S, int warp, int lane, float (&gh)[32 / NW][4]) {
    constexpr int R = 32 / NW;
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) gh[k][ci] = 0.f;
#pragma unroll 1
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            float4 w[4];
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) w[ci] = *reinterpret_cast<const float4*>(&S.P.w3[dy][dx][ci][0]);
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const float4 gp = S.g[(warp + k * NW - dy + 2) * 34 + (lane - dx + 2)];   // padded index of pixel (r-dy+1, c-dx+1)
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) gh[k][ci] += gp.x * w[ci].x + gp.y * w[ci].y + gp.z * w[ci].z + gp.w * w[ci].w;
            }
        }
}
// transposed conv-1: g_x0(r, c)[ci] = sum_{dy,dx,o} W1[dy][dx][ci][o] * g_c1(r-dy+1, c-dx+1)[o]
template <int NW>
__device__ __forceinline__ void td_conv1t_rows(const TdSmem& S, int warp, int lane, float (&gx0)[32 / NW][2]) {
    constexpr int R = 32 / NW;
#pragma unroll
    for (int k = 0; k < R; ++k) gx0[k][0] = gx0[k][1] = 0.f;
#pragma unroll 1
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const float4 w0 = *reinterpret_cast<const float4*>(&S.P.w1[dy][dx][0][0]);
            const float4 w1 = *reinterpret_cast<const float4*>(&S.P.w1[dy][dx][1][0]);
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const float4 g = S.g[(warp + k * NW - dy + 2) * 34 + (lane - dx + 2)];
                gx0[k][0] += g.x * w0.x + g.y * w0.y + g.z * w0.z + g.w * w0.w;
                gx0[k][1] += g.x * w1.x + g.y * w1.y + g.z * w1.z + g.w * w1.w;
            }
        }
}

// stats: double[16] of this coupling = sum c1[4], sum c1^2[4], sum c2[4], sum c2^2[4] over the batch.
// need_bn: how many of the two BatchNorms must be valid (0, 1, 2).
// reuse: the weights of this coupling are already in S.P (previous pass of the fused kernel): only the BatchNorm statistics in
// force and the accumulators are refreshed.
__device__ void td_load_params(TdSmem& S, const TdCoupling& d, const float* __restrict__ vars, const float* A,
                               const double* stats, double inv_cnt, int need_bn, bool reuse = false) {
    const int t = threadIdx.x;
    float* w1 = &S.P.w1[0][0][0][0];
    float* w2 = &S.P.w2[0][0];
    float* w3 = &S.P.w3[0][0][0][0];
    if (!reuse) {
        for (int k = t; k < 72; k += blockDim.x) w1[k] = vars[d.off_w1 + k];
        for (int k = t; k < 16; k += blockDim.x) w2[k] = vars[d.off_w2 + k];
        for (int k = t; k < 180; k += blockDim.x) w3[k] = vars[d.off_w3 + k];
        if (t < 16) (&S.P.A[0][0])[t] = d.has_mix ? __ldcg(A + t) : ((t >> 2) == (t & 3) ? 1.f : 0.f);
    }
    for (int k = t; k < (int)(blockDim.x >> 5) * NF_G_COUPLING_DOUBLES; k += blockDim.x) (&S.acc[0][0])[k] = 0.f;
    if (t < 4) {
        if (!reuse) {
            S.P.b1[t] = vars[d.off_b1 + t];
            S.P.b2[t] = vars[d.off_b2 + t];
            S.P.b3[t] = vars[d.off_b3 + t];
            S.P.logs[t] = vars[d.off_logs + t];
        }
        float m[2] = {0.f, 0.f}, is[2] = {1.f, 1.f};
        for (int j = 0; j < 2; ++j) {
            if (j >= need_bn) break;
            float mean, var;
            if (d.batch_stats) {
                const double mu = __ldcg(stats + 8 * j + t) * inv_cnt;       // written by other CTAs (atomics at L2)
                double v = __ldcg(stats + 8 * j + 4 + t) * inv_cnt - mu * mu;      // population variance (tf.nn.moments)
                if (v < 0.0) v = 0.0;
                mean = (float)mu;
                var = (float)v;
            } else {
                mean = vars[d.off_bn[2 * j] + t];
                var = vars[d.off_bn[2 * j + 1] + t];
            }
            m[j] = mean;
            is[j] = (float)(1.0 / sqrt((double)var + (double)d.bn_eps));
        }
        S.P.m1[t] = m[0]; S.P.is1[t] = is[0];
        S.P.m2[t] = m[1]; S.P.is2[t] = is[1];
    }
    if (t == 0) { S.P.scale = vars[d.off_scale]; S.P.has_mix = d.has_mix; }
}

template <int NW>
__device__ __forceinline__ void td_load_mixed(TdSmem& S, const float4* zin, int warp, int lane) {
    for (int r = warp; r < 32; r += NW) {
        float4 z = zin[r * 32 + lane];
        if (S.P.has_mix) z = mix_fwd(z, &S.P.A[0][0]);
        S.zp[r * 32 + lane] = z;
    }
}

// flush the CTA partial sums [lo, hi) with one fp64 atomic each
__device__ __forceinline__ void td_flush(TdSmem& S, double* dst, int lo, int hi) {
    __syncthreads();
    for (int k = lo + (int)threadIdx.x; k < hi; k += blockDim.x) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += S.acc[w][k];
        if (v != 0.f) atomicAdd(dst + k, (double)v);
    }
}

// ---------------------------------------------------------------------------------------------- forward
// STAGE 1: batch sums of c1 ; STAGE 2: batch sums of c2 ; STAGE 3: apply the coupling, accumulate the log-det.
// (The bodies are device functions shared by the per-pass kernels and the fused cooperative kernel td_step_kernel; buffers
// that the fused kernel rewrites while it runs are therefore NOT declared __restrict__ / read-only.)
template <int STAGE, int NW>
__device__ __forceinline__ void
td_fwd_body(TdSmem& S, const TdCoupling& d, const float* __restrict__ vars, const float* A, double* stats,
            const float4* zin, float4* zout, float* ld, long long n, double inv_cnt, bool reuse = false) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // reuse (fused kernel, one patch per CTA): weights and the mixed patch z' are still in shared memory from the previous stage
    td_load_params(S, d, vars, A, stats, inv_cnt, STAGE - 1, reuse);
    __syncthreads();
    float e3[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) e3[o] = expf(3.f * S.P.logs[o]);
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        if (!reuse) td_load_mixed<NW>(S, zin + p * NF_PIXELS, warp, lane);
        if (STAGE == 3)
            for (int k = threadIdx.x; k < 34 * 34; k += blockDim.x)
                if (on_ring(k)) S.h2[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        if (!reuse) td_conv1_rows<NW>(S, warp, lane);
        if (STAGE == 1) {
            float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
            for (int r = warp; r < 32; r += NW) {
                float c1[4];
                td_c1(S, r, lane, c1);
#pragma unroll
                for (int o = 0; o < 4; ++o) { s[o] += c1[o]; q[o] = fmaf(c1[o], c1[o], q[o]); }
            }
            { const float v8[8] = {s[0], s[1], s[2], s[3], q[0], q[1], q[2], q[3]}; cta_acc_vec<8>(S.acc[warp], v8, lane); }
        } else if (STAGE == 2) {
            float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
            for (int r = warp; r < 32; r += NW) {
                float c1[4], h1[4];
                td_c1(S, r, lane, c1);
#pragma unroll
                for (int o = 0; o < 4; ++o) h1[o] = fmaxf((c1[o] - S.P.m1[o]) * S.P.is1[o], 0.f);
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    float c2 = S.P.b2[o];
#pragma unroll
                    for (int i = 0; i < 4; ++i) c2 = fmaf(h1[i], S.P.w2[i][o], c2);
                    s[o] += c2;
                    q[o] = fmaf(c2, c2, q[o]);
                }
            }
            { const float v8[8] = {s[0], s[1], s[2], s[3], q[0], q[1], q[2], q[3]}; cta_acc_vec<8>(S.acc[warp] + 8, v8, lane); }
        } else {
            for (int r = warp; r < 32; r += NW) {
                float c1hat[4], h1[4], c2hat[4];
                td_net_to_c2hat(S, r, lane, c1hat, h1, c2hat);
                S.h2[(r + 1) * 34 + lane + 1] = make_float4(fmaxf(c2hat[0], 0.f), fmaxf(c2hat[1], 0.f), fmaxf(c2hat[2], 0.f), fmaxf(c2hat[3], 0.f));
            }
            __syncthreads();
            float lsum = 0.f;
            float pre_rows[32 / NW][4];
            td_conv3_rows<NW>(S, warp, lane, pre_rows);
#pragma unroll
            for (int k = 0; k < 32 / NW; ++k) {
                const int r = warp + k * NW;
                const float (&pre)[4] = pre_rows[k];
                const float sh0 = pre[0] * e3[0], sh1 = pre[1] * e3[1];
                const float ls0 = S.P.scale * tanhf(pre[2] * e3[2]), ls1 = S.P.scale * tanhf(pre[3] * e3[3]);
                const float4 zp = S.zp[r * 32 + lane];
                zout[p * NF_PIXELS + r * 32 + lane] = make_float4(zp.x, zp.y, fmaf(zp.z, expf(ls0), sh0), fmaf(zp.w, expf(ls1), sh1));
                lsum += ls0 + ls1;
            }
            lsum = tw_sum(lsum);                      // deterministic: fixed tree per warp, warps added in order
            if (lane == 0) S.red[warp] = lsum;
            __syncthreads();
            if (threadIdx.x == 0) {
                float tot = 0.f;
                for (int w = 0; w < NW; ++w) tot += S.red[w];
                if (ld) ld[p] += tot;
            }
        }
        __syncthreads();
    }
    if (STAGE == 1) td_flush(S, stats, 0, 8);
    if (STAGE == 2) td_flush(S, stats, 8, 16);
}
template <int STAGE, int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1)
td_fwd_kernel(const TdCoupling d, const float* __restrict__ vars, const float* A, double* stats,
              const float4* zin, float4* zout, float* ld, long long n, double inv_cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    td_fwd_body<STAGE, NW>(*reinterpret_cast<TdSmem*>(smem_raw), d, vars, A, stats, zin, zout, ld, n, inv_cnt);
}

// ---------------------------------------------------------------------------------------------- pass B1
// G_out -> g_shift, g_ls, g_x1 ; grads of scale, logs, b3, W3 ; g_h2 (transposed conv) ; BatchNorm-2 sums
template <int NW>
__device__ __forceinline__ void
td_b1_body(TdSmem& S, const TdCoupling& d, const float* __restrict__ vars, const float* A, const double* stats,
           const float4* zin, const float4* gout, float4* gzp, float4* scratch, long long n, float inv_n, double inv_cnt,
           double* grads) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    td_load_params(S, d, vars, A, stats, inv_cnt, 2);
    __syncthreads();
    float e3[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) e3[o] = expf(3.f * S.P.logs[o]);
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        td_load_mixed<NW>(S, zin + p * NF_PIXELS, warp, lane);
        for (int k = threadIdx.x; k < 34 * 34; k += blockDim.x)
            if (on_ring(k)) { S.h2[k] = make_float4(0.f, 0.f, 0.f, 0.f); S.g[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
        __syncthreads();
        td_conv1_rows<NW>(S, warp, lane);
        for (int r = warp; r < 32; r += NW) {
            float c1hat[4], h1[4], c2hat[4];
            td_net_to_c2hat(S, r, lane, c1hat, h1, c2hat);
            S.h2[(r + 1) * 34 + lane + 1] = make_float4(fmaxf(c2hat[0], 0.f), fmaxf(c2hat[1], 0.f), fmaxf(c2hat[2], 0.f), fmaxf(c2hat[3], 0.f));
        }
        __syncthreads();
        // conv-3 forward + coupling backward -> g_pre3 image, partial G_z'
        float g_scale = 0.f, g_logs[4] = {0.f, 0.f, 0.f, 0.f}, g_b3[4] = {0.f, 0.f, 0.f, 0.f};
        float pre_rows[32 / NW][4];
        td_conv3_rows<NW>(S, warp, lane, pre_rows);
#pragma unroll
        for (int k = 0; k < 32 / NW; ++k) {
            const int r = warp + k * NW;
            const float (&pre)[4] = pre_rows[k];
            float h3[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) h3[o] = pre[o] * e3[o];
            const float t0 = tanhf(h3[2]), t1 = tanhf(h3[3]);
            const float ls0 = S.P.scale * t0, ls1 = S.P.scale * t1, el0 = expf(ls0), el1 = expf(ls1);
            const float4 zp = S.zp[r * 32 + lane];
            const float4 go = gout[p * NF_PIXELS + r * 32 + lane];
            const float gls0 = go.z * zp.z * el0 - inv_n, gls1 = go.w * zp.w * el1 - inv_n;   // loss has -ldj/N
            g_scale += gls0 * t0 + gls1 * t1;
            const float gh3[4] = {go.z, go.w, gls0 * S.P.scale * (1.f - t0 * t0), gls1 * S.P.scale * (1.f - t1 * t1)};
            float gp[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) { g_logs[o] += 3.f * h3[o] * gh3[o]; gp[o] = gh3[o] * e3[o]; g_b3[o] += gp[o]; }
            S.g[(r + 1) * 34 + lane + 1] = make_float4(gp[0], gp[1], gp[2], gp[3]);
            gzp[p * NF_PIXELS + r * 32 + lane] = make_float4(go.x, go.y, go.z * el0, go.w * el1);   // x0 part completed in B3
        }
        __syncthreads();
        {   // slots NF_G_B3 .. NF_G_SCALE are consecutive: b3[4], logs[4], scale
            static_assert(NF_G_LOGS == NF_G_B3 + 4 && NF_G_SCALE == NF_G_B3 + 8, "gradient block layout");
            const float v9[9] = {g_b3[0], g_b3[1], g_b3[2], g_b3[3], g_logs[0], g_logs[1], g_logs[2], g_logs[3], g_scale};
            cta_acc_vec<9>(S.acc[warp] + NF_G_B3, v9, lane);
        }
        // grad W3[dy][dx][ci][o] = sum_pixels in(r+dy, c+dx)[ci] * g_pre3(r, c)[o]   (ci = 4: ring indicator)
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                float a[20];            // [ci][o], the 20 consecutive slots of this tap
#pragma unroll
                for (int k = 0; k < 20; ++k) a[k] = 0.f;
                for (int r = warp; r < 32; r += NW) {
                    const int R = r + dy, C = lane + dx;
                    const float4 h = S.h2[R * 34 + C];
                    const float ring = (R == 0 || R == 33 || C == 0 || C == 33) ? 1.f : 0.f;
                    const float4 gp = S.g[(r + 1) * 34 + lane + 1];
                    const float hv[5] = {h.x, h.y, h.z, h.w, ring}, gv[4] = {gp.x, gp.y, gp.z, gp.w};
#pragma unroll
                    for (int ci = 0; ci < 5; ++ci)
#pragma unroll
                        for (int o = 0; o < 4; ++o) a[ci * 4 + o] = fmaf(hv[ci], gv[o], a[ci * 4 + o]);
                }
                cta_acc_vec<20>(S.acc[warp] + NF_G_W3 + (dy * 3 + dx) * 20, a, lane);
            }
        // g_h2(r, c)[ci] = sum_{dy,dx,o} W3[dy][dx][ci][o] * g_pre3(r-dy+1, c-dx+1)[o] ; ReLU mask ; BatchNorm-2 sums
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        float gh_rows[32 / NW][4];
        td_conv3t_rows<NW>(S, warp, lane, gh_rows);
#pragma unroll
        for (int k = 0; k < 32 / NW; ++k) {
            const int r = warp + k * NW;
            const float (&gh)[4] = gh_rows[k];
            const float4 h = S.h2[(r + 1) * 34 + lane + 1];
            float gc[4];
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
                const float hv = comp(h, ci);
                gc[ci] = hv > 0.f ? gh[ci] : 0.f;
                s1[ci] += gc[ci];
                s2[ci] += gc[ci] * hv;          // = g_c2hat * c2hat wherever the mask is on
            }
            scratch[p * NF_PIXELS + r * 32 + lane] = make_float4(gc[0], gc[1], gc[2], gc[3]);
        }
        { const float v8[8] = {s1[0], s1[1], s1[2], s1[3], s2[0], s2[1], s2[2], s2[3]}; cta_acc_vec<8>(S.acc[warp] + NF_G_BN2, v8, lane); }
        __syncthreads();
    }
    td_flush(S, grads, 0, NF_G_COUPLING_DOUBLES);
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1)
td_b1_kernel(const TdCoupling d, const float* __restrict__ vars, const float* A, const double* stats, const float4* zin,
             const float4* gout, float4* gzp, float4* scratch, long long n, float inv_n, double inv_cnt, double* grads) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    td_b1_body<NW>(*reinterpret_cast<TdSmem*>(smem_raw), d, vars, A, stats, zin, gout, gzp, scratch, n, inv_n, inv_cnt, grads);
}

// ---------------------------------------------------------------------------------------------- pass B2
// BatchNorm-2 backward ; grads of W2, b2 ; g_h1 ; BatchNorm-1 sums
template <int NW>
__device__ __forceinline__ void
td_b2_body(TdSmem& S, const TdCoupling& d, const float* __restrict__ vars, const float* A, const double* stats,
           const float4* zin, float4* scratch, long long n, double inv_cnt, double* grads, bool reuse = false) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float bn2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bn2[k] = d.batch_stats ? (float)(__ldcg(grads + NF_G_BN2 + k) * inv_cnt) : 0.f;   // other CTAs' atomics
    td_load_params(S, d, vars, A, stats, inv_cnt, 2, reuse);
    __syncthreads();
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        if (!reuse) td_load_mixed<NW>(S, zin + p * NF_PIXELS, warp, lane);
        __syncthreads();
        if (!reuse) td_conv1_rows<NW>(S, warp, lane);
        float gw2[4][4], gb2[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int o = 0; o < 4; ++o) gw2[i][o] = 0.f;
        for (int r = warp; r < 32; r += NW) {
            float c1hat[4], h1[4], c2hat[4];
            td_net_to_c2hat(S, r, lane, c1hat, h1, c2hat);
            const float4 gc4 = scratch[p * NF_PIXELS + r * 32 + lane];
            float gc2[4], gh1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                gc2[o] = (comp(gc4, o) - bn2[o] - c2hat[o] * bn2[4 + o]) * S.P.is2[o];
                gb2[o] += gc2[o];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int o = 0; o < 4; ++o) { gw2[i][o] = fmaf(h1[i], gc2[o], gw2[i][o]); gh1[i] = fmaf(gc2[o], S.P.w2[i][o], gh1[i]); }
            float gc1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                gc1[i] = h1[i] > 0.f ? gh1[i] : 0.f;
                t1[i] += gc1[i];
                t2[i] += gc1[i] * h1[i];
            }
            scratch[p * NF_PIXELS + r * 32 + lane] = make_float4(gc1[0], gc1[1], gc1[2], gc1[3]);
        }
        {   // W2[4][4] and b2[4] are 20 consecutive slots; the BatchNorm-1 sums 8 more
            static_assert(NF_G_B2 == NF_G_W2 + 16, "gradient block layout");
            float v20[20];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int o = 0; o < 4; ++o) v20[i * 4 + o] = gw2[i][o];
                v20[16 + i] = gb2[i];
            }
            cta_acc_vec<20>(S.acc[warp] + NF_G_W2, v20, lane);
            const float v8[8] = {t1[0], t1[1], t1[2], t1[3], t2[0], t2[1], t2[2], t2[3]};
            cta_acc_vec<8>(S.acc[warp] + NF_G_BN1, v8, lane);
        }
        __syncthreads();
    }
    td_flush(S, grads, 0, NF_G_COUPLING_DOUBLES);
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1)
td_b2_kernel(const TdCoupling d, const float* __restrict__ vars, const float* A, const double* stats, const float4* zin,
             float4* scratch, long long n, double inv_cnt, double* grads) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    td_b2_body<NW>(*reinterpret_cast<TdSmem*>(smem_raw), d, vars, A, stats, zin, scratch, n, inv_cnt, grads);
}

// ---------------------------------------------------------------------------------------------- pass B3
// BatchNorm-1 backward ; grads of W1, b1 ; g_x0 (transposed conv) ; grad of A ; G_in = g_z' . A^T
template <int NW>
__device__ __forceinline__ void
td_b3_body(TdSmem& S, const TdCoupling& d, const float* __restrict__ vars, const float* A, const double* stats,
           const float4* zin, const float4* scratch, const float4* gzp, float4* gin, long long n, double inv_cnt, double* grads,
           bool reuse = false) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float bn1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bn1[k] = d.batch_stats ? (float)(__ldcg(grads + NF_G_BN1 + k) * inv_cnt) : 0.f;
    td_load_params(S, d, vars, A, stats, inv_cnt, 2, reuse);
    __syncthreads();
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        if (!reuse) td_load_mixed<NW>(S, zin + p * NF_PIXELS, warp, lane);
        for (int k = threadIdx.x; k < 34 * 34; k += blockDim.x)
            if (on_ring(k)) S.g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        if (!reuse) td_conv1_rows<NW>(S, warp, lane);
        float gb1[4] = {0.f, 0.f, 0.f, 0.f};
        for (int r = warp; r < 32; r += NW) {
            float c1[4];
            td_c1(S, r, lane, c1);
            const float4 g4 = scratch[p * NF_PIXELS + r * 32 + lane];
            float gc1[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const float c1hat = (c1[o] - S.P.m1[o]) * S.P.is1[o];
                gc1[o] = (comp(g4, o) - bn1[o] - c1hat * bn1[4 + o]) * S.P.is1[o];
                gb1[o] += gc1[o];
            }
            S.g[(r + 1) * 34 + lane + 1] = make_float4(gc1[0], gc1[1], gc1[2], gc1[3]);
        }
        __syncthreads();
        cta_acc_vec<4>(S.acc[warp] + NF_G_B1, gb1, lane);
        // grad W1[dy][dx][ci][o] = sum_pixels x0(r+dy-1, c+dx-1)[ci] * g_c1(r, c)[o]
        for (int dy = 0; dy < 3; ++dy) {
            float a[24];                // [dx][ci][o]: the 24 consecutive slots of filter row dy
#pragma unroll
            for (int k = 0; k < 24; ++k) a[k] = 0.f;
            for (int r = warp; r < 32; r += NW) {
                const int rr = r + dy - 1;
                if (rr < 0 || rr > 31) continue;
                const float4 g = S.g[(r + 1) * 34 + lane + 1];
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int cc = lane + dx - 1;
                    if (cc < 0 || cc > 31) continue;
                    const float4 z = S.zp[rr * 32 + cc];
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        a[dx * 8 + o] = fmaf(z.x, comp(g, o), a[dx * 8 + o]);
                        a[dx * 8 + 4 + o] = fmaf(z.y, comp(g, o), a[dx * 8 + 4 + o]);
                    }
                }
            }
            cta_acc_vec<24>(S.acc[warp] + NF_G_W1 + dy * 24, a, lane);
        }
        // g_x0 (transposed conv), complete g_z', grad A, G_in
        float gA[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int o = 0; o < 4; ++o) gA[i][o] = 0.f;
        float gx0_rows[32 / NW][2];
        td_conv1t_rows<NW>(S, warp, lane, gx0_rows);
#pragma unroll
        for (int k = 0; k < 32 / NW; ++k) {
            const int r = warp + k * NW;
            const float (&gx0)[2] = gx0_rows[k];
            float4 gz = gzp[p * NF_PIXELS + r * 32 + lane];
            gz.x += gx0[0];
            gz.y += gx0[1];
            float4 out = gz;
            if (S.P.has_mix) {
                const float4 zi = zin[p * NF_PIXELS + r * 32 + lane];
                const float zv[4] = {zi.x, zi.y, zi.z, zi.w}, gv[4] = {gz.x, gz.y, gz.z, gz.w};
                float gi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int o = 0; o < 4; ++o) { gA[i][o] = fmaf(zv[i], gv[o], gA[i][o]); gi[i] = fmaf(gv[o], S.P.A[i][o], gi[i]); }
                out = make_float4(gi[0], gi[1], gi[2], gi[3]);
            }
            gin[p * NF_PIXELS + r * 32 + lane] = out;
        }
        if (S.P.has_mix) {
            float v16[16];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int o = 0; o < 4; ++o) v16[i * 4 + o] = gA[i][o];
            cta_acc_vec<16>(S.acc[warp] + NF_G_A, v16, lane);
        }
        __syncthreads();
    }
    td_flush(S, grads, 0, NF_G_COUPLING_DOUBLES);
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1)
td_b3_kernel(const TdCoupling d, const float* __restrict__ vars, const float* A, const double* stats, const float4* zin,
             const float4* scratch, const float4* gzp, float4* gin, long long n, double inv_cnt, double* grads) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    td_b3_body<NW>(*reinterpret_cast<TdSmem*>(smem_raw), d, vars, A, stats, zin, scratch, gzp, gin, n, inv_cnt, grads);
}

// ---------------------------------------------------------------------------------------------- scale layers
// deterministic CTA sum (fixed shuffle tree per warp, warps added in order by thread 0); result valid in thread 0
__device__ __forceinline__ float cta_sum_det(float v, float* sh, int warp, int lane) {
    v = tw_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float tot = 0.f;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sh[w];
    return tot;
}

// table: [NF_MAX_ROWS][2] = (a, b) for sdn (scale^2 = a*y + b), (g, -) for gain.  z_out = z_in / scale.
__device__ __forceinline__ void
td_scale_fwd_body(float* sh, const float* table, int is_sdn, int full_sum, const float4* zin, const float4* __restrict__ y,
                  float4* zout, float* ld, const int* __restrict__ rows, int default_row, long long n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        int row = rows ? rows[p] : default_row;
        row = min(max(row, 0), NF_MAX_ROWS - 1);
        const float a = __ldcg(table + row * 2), b = __ldcg(table + row * 2 + 1);
        float lsum = 0.f;
        for (int k = threadIdx.x; k < NF_PIXELS; k += blockDim.x) {
            const long long idx = p * NF_PIXELS + k;
            const float4 z = zin[idx];
            if (is_sdn) {
                const float4 yv = y[idx];
                const float v0 = fmaf(a, yv.x, b), v1 = fmaf(a, yv.y, b), v2 = fmaf(a, yv.z, b), v3 = fmaf(a, yv.w, b);
                zout[idx] = make_float4(z.x * rsqrtf(v0), z.y * rsqrtf(v1), z.z * rsqrtf(v2), z.w * rsqrtf(v3));
                lsum -= 0.5f * (logf(v0) + logf(v1) + logf(v2) + logf(v3));       // -sum log scale
            } else {
                const float gi = 1.f / a;
                zout[idx] = make_float4(z.x * gi, z.y * gi, z.z * gi, z.w * gi);
            }
        }
        const float tot = cta_sum_det(lsum, sh, warp, lane);
        if (threadIdx.x == 0 && ld) ld[p] += is_sdn ? tot : -(full_sum ? (float)NF_DIMS : 1.f) * logf(a);
    }
}
__global__ void __launch_bounds__(TD_THREADS)
td_scale_fwd_kernel(const float* table, int is_sdn, int full_sum, const float4* zin, const float4* __restrict__ y, float4* zout,
                    float* ld, const int* __restrict__ rows, int default_row, long long n) {
    __shared__ float sh[16];
    td_scale_fwd_body(sh, table, is_sdn, full_sum, zin, y, zout, ld, rows, default_row, n);
}

// G_in = G_out / scale ; table-row gradients (see nf_train_scale_kernel in nf_train.cu)
__device__ __forceinline__ void
td_scale_bwd_body(float* sh, const float* table, int is_sdn, int full_sum, const float4* zout, const float4* __restrict__ y,
                  const float4* gout, float4* gin, const int* __restrict__ rows, int default_row, long long n, float inv_n,
                  double* grads) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        int row = rows ? rows[p] : default_row;
        row = min(max(row, 0), NF_MAX_ROWS - 1);
        const float a = __ldcg(table + row * 2), b = __ldcg(table + row * 2 + 1);
        float ga = 0.f, gb = 0.f;
        for (int k = threadIdx.x; k < NF_PIXELS; k += blockDim.x) {
            const long long idx = p * NF_PIXELS + k;
            const float4 zo = zout[idx], go = gout[idx];
            if (is_sdn) {
                const float4 yv = y[idx];
                const float yy[4] = {yv.x, yv.y, yv.z, yv.w}, zz[4] = {zo.x, zo.y, zo.z, zo.w}, gg[4] = {go.x, go.y, go.z, go.w};
                float o4[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float v = fmaf(a, yy[c], b), rv = rsqrtf(v);
                    const float dv = (-0.5f * gg[c] * zz[c] + 0.5f * inv_n) / v;
                    ga += dv * yy[c];
                    gb += dv;
                    o4[c] = gg[c] * rv;
                }
                gin[idx] = make_float4(o4[0], o4[1], o4[2], o4[3]);
            } else {
                const float ginv = 1.f / a;
                ga += -(go.x * zo.x + go.y * zo.y + go.z * zo.z + go.w * zo.w) * ginv;
                gin[idx] = make_float4(go.x * ginv, go.y * ginv, go.z * ginv, go.w * ginv);
            }
        }
        const float ta = cta_sum_det(ga, sh, warp, lane);
        const float tb = cta_sum_det(gb, sh, warp, lane);
        if (threadIdx.x == 0) {
            float fa = ta;
            if (!is_sdn) fa += (full_sum ? (float)NF_DIMS : 1.f) * inv_n / a;
            atomicAdd(grads + row * 2, (double)fa);
            if (is_sdn) atomicAdd(grads + row * 2 + 1, (double)tb);
        }
    }
}
__global__ void __launch_bounds__(TD_THREADS)
td_scale_bwd_kernel(const float* table, int is_sdn, int full_sum, const float4* zout, const float4* __restrict__ y,
                    const float4* gout, float4* gin, const int* __restrict__ rows, int default_row, long long n, float inv_n,
                    double* grads) {
    __shared__ float sh[16];
    td_scale_bwd_body(sh, table, is_sdn, full_sum, zout, y, gout, gin, rows, default_row, n, inv_n, grads);
}

// prior + loss terms of one patch, and the seed of the backward sweep G = z / N:
//   nll = -(ldj + ldj_const + sum -0.5 (log 2pi + z^2)) ; sd_z = sqrt(var_hwc(z))   (noise_flow_model.py:458-480)
__device__ __forceinline__ void
td_nll_body(float* sh, const float4* z, const float* ld, const double* consts, float4* g, float* nll, float* sdz, long long n,
            float inv_n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        float s1 = 0.f, s2 = 0.f;
        for (int k = threadIdx.x; k < NF_PIXELS; k += blockDim.x) {
            const float4 v = z[p * NF_PIXELS + k];
            s1 += v.x + v.y + v.z + v.w;
            s2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            if (g) g[p * NF_PIXELS + k] = make_float4(v.x * inv_n, v.y * inv_n, v.z * inv_n, v.w * inv_n);
        }
        const float t1 = cta_sum_det(s1, sh, warp, lane);
        const float t2 = cta_sum_det(s2, sh, warp, lane);
        if (threadIdx.x == 0) {
            const double S1 = t1, S2 = t2;
            const double logp = -0.5 * ((double)NF_DIMS * 1.8378770664093453 + S2);
            if (nll) nll[p] = (float)(-((double)ld[p] + __ldcg(consts) + logp));
            const double mean = S1 / NF_DIMS;
            double var = S2 / NF_DIMS - mean * mean;
            if (var < 0.0) var = 0.0;
            if (sdz) sdz[p] = (float)sqrt(var);
        }
    }
}
__global__ void __launch_bounds__(TD_THREADS)
td_nll_kernel(const float4* z, const float* ld, const double* consts, float4* g, float* nll, float* sdz, long long n, float inv_n) {
    __shared__ float sh[16];
    td_nll_body(sh, z, ld, consts, g, nll, sdz, n, inv_n);
}

// ---------------------------------------------------------------------------------------------- small device-side host work
struct TdOp {                    // device copy of nf_train_op plus derived indices
    nf_train_op o;
    int32_t cidx;                // coupling index (statistics / gradient block) or -1
    int32_t sidx;                // scale-layer index (table / row-gradient block) or -1
};
struct TdProgram {
    int32_t n_ops;
    int32_t tri_lo[6], tri_up[6];
    int32_t pad_;
    TdOp ops[TD_MAX_OPS];
};

__device__ const double kIsoVals[5] = {100.0, 400.0, 800.0, 1600.0, 3200.0};   // cond_utils.py:224

// L (unit lower), U' = U + diag(sign_S exp(log_S)) from the LU variables (matrix_param.py:117-130)
__device__ void td_build_lu(const nf_train_op& o, const int* tri_lo, const int* tri_up, const float* vars, double (&P)[4][4],
                            double (&L)[4][4], double (&U)[4][4], double (&sd)[4]) {
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { P[r][c] = vars[o.off_P + r * 4 + c]; L[r][c] = r == c ? 1.0 : 0.0; U[r][c] = 0.0; }
    for (int k = 0; k < 6; ++k) {
        L[tri_lo[k] >> 2][tri_lo[k] & 3] = vars[o.off_L + k];
        U[tri_up[k] >> 2][tri_up[k] & 3] = vars[o.off_U + k];
    }
    for (int j = 0; j < 4; ++j) { sd[j] = (double)vars[o.off_signS + j] * exp((double)vars[o.off_logS + j]); U[j][j] = sd[j]; }
}

// (a, b) of a scale row: cond_utils.py:165-187 (sdn4), :205-239 (sdn5), :242-276 (sdn6), :432-440 (gain4)
__device__ void td_scale_row(const nf_train_op& o, const float* vars, int cam, int isoi, double& a, double& b) {
    const double iso = kIsoVals[isoi], c = o.c_i;
    if (o.token == NF_TOKEN_GAIN4) { a = vars[o.off_gain_val]; b = 0.0; return; }
    const double g = vars[o.off_gain_params + isoi], beta1 = vars[o.off_beta1], beta2 = vars[o.off_beta2];
    if (o.token == NF_TOKEN_SDN4) {
        a = exp(beta1) / (exp(g) * iso);
        b = exp(beta2);
    } else if (o.token == NF_TOKEN_SDN5) {
        const double o0 = exp(c * vars[o.off_cam_params + cam]), o1 = exp(c * vars[o.off_cam_params + 5 + cam]),
                     o2 = exp(c * vars[o.off_cam_params + 10 + cam]);
        a = exp(c * beta1 * o0) / (exp(c * g * o2) * iso);
        b = exp(c * beta2 * o1);
    } else {   // sdn6
        const double o0 = exp(c * vars[o.off_cam_params + cam]);
        a = exp(c * beta1) / (exp(c * g * o0) * iso);
        b = exp(c * beta2);
    }
}

// derived 1x1 matrices, scale tables, constant log-det of the chain.  Work items = (op, table row): a coupling's matrix is
// item (op, 0); the 32 rows of a scale table (six fp64 exp each for sdn5) go to 32 threads instead of one.
__device__ __forceinline__ void td_prep_body(const TdProgram* __restrict__ prog, const float* __restrict__ vars, float* Amat,
                                             float* tables, double* consts) {
  for (int item = threadIdx.x; item < prog->n_ops * NF_MAX_ROWS; item += blockDim.x) {
    const int i = item / NF_MAX_ROWS, row = item % NF_MAX_ROWS;
    const TdOp& op = prog->ops[i];
    const nf_train_op& o = op.o;
    if (o.kind == NF_TOP_COUPLING) {
        if (row != 0) continue;
        float* A = Amat + op.cidx * 16;
        if (o.mix_kind == 1) {
            double P[4][4], L[4][4], U[4][4], sd[4], LU[4][4];
            td_build_lu(o, prog->tri_lo, prog->tri_up, vars, P, L, U, sd);
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) { double s = 0.0; for (int k = 0; k < 4; ++k) s += L[r][k] * U[k][c]; LU[r][c] = s; }
            double ls = 0.0;
            for (int r = 0; r < 4; ++r) {
                ls += (double)vars[o.off_logS + r];
                for (int c = 0; c < 4; ++c) { double s = 0.0; for (int k = 0; k < 4; ++k) s += P[r][k] * LU[k][c]; A[r * 4 + c] = (float)s; }
            }
            atomicAdd(consts, (double)NF_PIXELS * ls);                       // layers.py:129-130
        } else if (o.mix_kind == 2) {
            for (int k = 0; k < 16; ++k) A[k] = 0.f;
            for (int k = 0; k < 4; ++k) A[k * 4 + o.perm[k]] = 1.f;          // inverse: out[perm[i]] = in[i]
        }
    } else if (o.kind == NF_TOP_SCALE) {
        float* T = tables + op.sidx * NF_MAX_ROWS * 2;
        double a = 1.0, b = 1.0;
        if (row < 25) td_scale_row(o, vars, row / 5, row % 5, a, b);
        T[row * 2] = (float)a;
        T[row * 2 + 1] = (float)b;
    }
  }
}
__global__ void __launch_bounds__(256) td_prep_kernel(const TdProgram* __restrict__ prog, const float* __restrict__ vars, float* Amat, float* tables,
                               double* consts) {
    td_prep_body(prog, vars, Amat, tables, consts);
}

// Chain rules from the kernel-level gradients to the TF variables, all into the flat reduce buffer:
//   red[0 .. n_vars)              d loss / d variable
//   red[n_vars + 3 + 16 c + j]    batch statistics of coupling c (mean1, var1, mean2, var2) for the moving averages
__global__ void __launch_bounds__(512) td_chain_kernel(const TdProgram* __restrict__ prog, const float* __restrict__ vars, const double* __restrict__ cgrads,
                                const double* __restrict__ sgrads, const double* __restrict__ stats, double inv_cnt,
                                int batch_stats, long long n_vars, double* __restrict__ red) {
    const int n_ops = prog->n_ops;
    // (1) coupling tensors: straight copies (checkpoint layouts are the kernel layouts); work items = (op, slot of its gradient
    //     block), all independent, so the loads of every coupling are in flight at once
    for (int item = threadIdx.x; item < n_ops * NF_G_COUPLING_DOUBLES; item += blockDim.x) {
        const int i = item / NF_G_COUPLING_DOUBLES, k = item % NF_G_COUPLING_DOUBLES;
        const TdOp& op = prog->ops[i];
        if (op.o.kind != NF_TOP_COUPLING) continue;
        const double* g = cgrads + (size_t)op.cidx * NF_G_COUPLING_DOUBLES;
        const nf_train_op& o = op.o;
        if (k < NF_G_HOST_COUPLING) {
            const int src = NF_G_W1 + k;
            int dst;
            if (src < NF_G_B1) dst = o.off_w1 + (src - NF_G_W1);
            else if (src < NF_G_W2) dst = o.off_b1 + (src - NF_G_B1);
            else if (src < NF_G_B2) dst = o.off_w2 + (src - NF_G_W2);
            else if (src < NF_G_W3) dst = o.off_b2 + (src - NF_G_B2);
            else if (src < NF_G_B3) dst = o.off_w3 + (src - NF_G_W3);
            else if (src < NF_G_LOGS) dst = o.off_b3 + (src - NF_G_B3);
            else if (src < NF_G_SCALE) dst = o.off_logs + (src - NF_G_LOGS);
            else dst = o.off_scale;
            atomicAdd(red + dst, g[src]);
        } else if (k < NF_G_HOST_COUPLING + 16) {
            const int j = k - NF_G_HOST_COUPLING, st = j >> 3, c = j & 3;   // j: [mean1 4][var1 4][mean2 4][var2 4]
            double val = 0.0;
            if (batch_stats) {
                const double mu = stats[op.cidx * 16 + 8 * st + c] * inv_cnt;
                if ((j & 4) == 0) val = (double)(float)mu;
                else {
                    double v = stats[op.cidx * 16 + 8 * st + 4 + c] * inv_cnt - mu * mu;
                    val = (double)(float)(v < 0.0 ? 0.0 : v);
                }
            }
            red[n_vars + 3 + op.cidx * 16 + j] = val;
        }
    }
    // (2) work items (op, table row): the LU chain rule (train.lu_chain) of a coupling is item (op, 0); the 25 rows of a scale
    //     layer chain to its variables independently (a loop over rows in one thread serialised 25 dependent L2 round trips)
    for (int item = threadIdx.x; item < n_ops * NF_MAX_ROWS; item += blockDim.x) {
    const int i = item / NF_MAX_ROWS, row = item % NF_MAX_ROWS;
    const TdOp& op = prog->ops[i];
    const nf_train_op& o = op.o;
    if (o.kind == NF_TOP_COUPLING && o.mix_kind == 1 && row == 0) {
        double P[4][4], L[4][4], U[4][4], sd[4], G[4][4], ptg[4][4], dL[4][4], dU[4][4];
        td_build_lu(o, prog->tri_lo, prog->tri_up, vars, P, L, U, sd);
        const double* g = cgrads + (size_t)op.cidx * NF_G_COUPLING_DOUBLES + NF_G_A;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) G[r][c] = g[r * 4 + c];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) { double s = 0.0; for (int k = 0; k < 4; ++k) s += P[k][r] * G[k][c]; ptg[r][c] = s; }      // P^T G
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                double s = 0.0, u = 0.0;
                for (int k = 0; k < 4; ++k) { s += ptg[r][k] * U[c][k]; u += L[k][r] * ptg[k][c]; }   // ptg U'^T ; L^T ptg
                dL[r][c] = s;
                dU[r][c] = u;
            }
        for (int k = 0; k < 6; ++k) {
            atomicAdd(red + o.off_L + k, dL[prog->tri_lo[k] >> 2][prog->tri_lo[k] & 3]);
            atomicAdd(red + o.off_U + k, dU[prog->tri_up[k] >> 2][prog->tri_up[k] & 3]);
        }
        // the layer's own log-det, H*W*sum(log_S) per patch, enters the loss as -ldj
        for (int j = 0; j < 4; ++j) atomicAdd(red + o.off_logS + j, dU[j][j] * sd[j] - (double)NF_PIXELS);
    } else if (o.kind == NF_TOP_SCALE) {
        const double* sg = sgrads + (size_t)op.sidx * NF_MAX_ROWS * 2;
        const double c = o.c_i;
        if (row < 25) {
            const double da = sg[row * 2], db = sg[row * 2 + 1];
            if (da == 0.0 && db == 0.0) continue;
            const int cam = row / 5, isoi = row % 5;
            if (o.token == NF_TOKEN_GAIN4) { atomicAdd(red + o.off_gain_val, da); continue; }
            double a, b;
            td_scale_row(o, vars, cam, isoi, a, b);
            const double g = vars[o.off_gain_params + isoi], beta1 = vars[o.off_beta1], beta2 = vars[o.off_beta2];
            if (o.token == NF_TOKEN_SDN4) {
                atomicAdd(red + o.off_beta1, da * a);
                atomicAdd(red + o.off_gain_params + isoi, -da * a);
                atomicAdd(red + o.off_beta2, db * b);
            } else if (o.token == NF_TOKEN_SDN5) {
                const double cp0 = vars[o.off_cam_params + cam], cp1 = vars[o.off_cam_params + 5 + cam], cp2 = vars[o.off_cam_params + 10 + cam];
                const double o0 = exp(c * cp0), o1 = exp(c * cp1), o2 = exp(c * cp2);
                atomicAdd(red + o.off_beta1, da * a * c * o0);
                atomicAdd(red + o.off_gain_params + isoi, -da * a * c * o2);
                atomicAdd(red + o.off_beta2, db * b * c * o1);
                atomicAdd(red + o.off_cam_params + cam, da * a * c * beta1 * o0 * c);
                atomicAdd(red + o.off_cam_params + 5 + cam, db * b * c * beta2 * o1 * c);
                atomicAdd(red + o.off_cam_params + 10 + cam, -da * a * c * g * o2 * c);
            } else {   // sdn6
                const double o0 = exp(c * vars[o.off_cam_params + cam]);
                atomicAdd(red + o.off_beta1, da * a * c);
                atomicAdd(red + o.off_gain_params + isoi, -da * a * c * o0);
                atomicAdd(red + o.off_beta2, db * b * c);
                atomicAdd(red + o.off_cam_params + cam, -da * a * c * g * o0 * c);
            }
        }
    }
    }
}

// Adam with TensorFlow's update rule + BatchNorm moving averages; one CTA so that the step counter has one writer
__global__ void __launch_bounds__(1024)
td_apply_kernel(const TdProgram* __restrict__ prog, float* __restrict__ vars, const unsigned char* __restrict__ trainable,
                double* __restrict__ am, double* __restrict__ av, long long* __restrict__ step, const double* __restrict__ red,
                long long n_vars, double lr, double b1, double b2, double eps, double inv_world, int update_bn) {
    const long long t = step[0] + 1;
    __shared__ double lr_t_s;
    if (threadIdx.x == 0) lr_t_s = lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t));
    __syncthreads();
    const double lr_t = lr_t_s;
    for (long long k = threadIdx.x; k < n_vars; k += blockDim.x) {
        if (!trainable[k]) continue;
        const double g = red[k] * inv_world;
        const double m = am[k] + (1.0 - b1) * (g - am[k]);
        const double v = av[k] + (1.0 - b2) * (g * g - av[k]);
        am[k] = m;
        av[k] = v;
        vars[k] = (float)((double)vars[k] - lr_t * m / (sqrt(v) + eps));
    }
    if (update_bn) {
        for (int item = threadIdx.x; item < prog->n_ops * 16; item += blockDim.x) {     // (op, statistic): all independent
            const TdOp& op = prog->ops[item >> 4];
            if (op.o.kind != NF_TOP_COUPLING) continue;
            const int j = item & 15;
            const int offs[4] = {op.o.off_bn1_mean, op.o.off_bn1_var, op.o.off_bn2_mean, op.o.off_bn2_var};
            const int dst = offs[j >> 2] + (j & 3);
            const float batch = (float)(red[n_vars + 3 + op.cidx * 16 + j] * inv_world);
            const float cur = vars[dst];
            vars[dst] = cur - TD_BN_MOMENTUM * (cur - batch);                 // layers.py:394-395
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) step[0] = t;
}

// ---------------------------------------------------------------------------------------------- the fused step
// Loss + gradient of one batch as ONE cooperative kernel.  At the reference's batch sizes (138-207 patches per GPU) the
// per-pass kernels above each run a single wave of one-patch CTAs for 9-27 us, and the 57 dependent launches of a step cost
// ~0.6 ms whatever the batch (0.63 ms at 64 patches, 0.64 ms at 138): the step is bound by launch + prologue latency, not
// by work.  Here every CTA keeps ITS patch for the whole step and walks the op list itself -- prep, forward (three stages
// per coupling), prior / NLL, backward (three passes per coupling) -- with a grid-wide barrier only where the math needs
// the whole batch: after prep (matrices / tables are built by CTA 0) and around the BatchNorm batch sums (2 per coupling and
// direction).  Activations and gradients still go through the global work space (a CTA re-reads only what it wrote
// itself: L1/L2 hits), so the pass bodies are shared with the per-pass kernels, which remain the path for batches larger
// than the number of co-resident CTAs and the cross-check of this kernel (tests/test_gpu_trainer.py).
struct TdStepArgs {
    const TdProgram* prog;
    const float* vars;
    float *Amat, *tables;
    double *stats, *cgrads, *sgrads, *consts;
    float* ws;                  // activation / gradient slots, `stride` floats each: [0, n_ops) op outputs, then gA gB scratch gzp, ld nll sdz
    long long stride;
    const float *x, *y;
    const int* rows;
    int default_row;
    long long n;
    int batch_stats;
    float bn_eps;
};

__device__ __forceinline__ TdCoupling td_desc(const nf_train_op& o, int batch_stats, float bn_eps) {
    TdCoupling d;
    d.off_w1 = o.off_w1; d.off_b1 = o.off_b1; d.off_w2 = o.off_w2; d.off_b2 = o.off_b2;
    d.off_w3 = o.off_w3; d.off_b3 = o.off_b3; d.off_logs = o.off_logs; d.off_scale = o.off_scale;
    d.off_bn[0] = o.off_bn1_mean; d.off_bn[1] = o.off_bn1_var; d.off_bn[2] = o.off_bn2_mean; d.off_bn[3] = o.off_bn2_var;
    d.has_mix = o.mix_kind != 0;
    d.batch_stats = batch_stats;
    d.bn_eps = bn_eps;
    d.pad_ = 0;
    return d;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1)
td_step_kernel(const TdStepArgs a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TdSmem& S = *reinterpret_cast<TdSmem*>(smem_raw);
    __shared__ float sh[16];
    const TdProgram* prog = a.prog;
    const int G = prog->n_ops;
    const long long n = a.n;
    const double inv_cnt = 1.0 / ((double)n * NF_PIXELS);
    const float inv_n = 1.f / (float)n;
    auto slot = [&](int k) { return a.ws + (long long)k * a.stride; };
    float *gA = slot(G), *gB = slot(G + 1), *scratch = slot(G + 2), *gzp = slot(G + 3);
    float *ld = slot(G + 4), *nll = ld + n, *sdz = nll + n;

    if (blockIdx.x == 0) td_prep_body(prog, a.vars, a.Amat, a.tables, a.consts);
    grid.sync();
    // ---- forward, keeping every op's input
    for (int i = 0; i < G; ++i) {
        const TdOp& op = prog->ops[i];
        const float4* in = (const float4*)(i == 0 ? a.x : slot(i - 1));
        float4* out = (float4*)slot(i);
        if (op.o.kind == NF_TOP_COUPLING) {
            const TdCoupling d = td_desc(op.o, a.batch_stats, a.bn_eps);
            const float* A = a.Amat + op.cidx * 16;
            double* st = a.stats + op.cidx * 16;
            if (a.batch_stats) {
                td_fwd_body<1, NW>(S, d, a.vars, A, st, in, out, ld, n, inv_cnt);
                grid.sync();
                td_fwd_body<2, NW>(S, d, a.vars, A, st, in, out, ld, n, inv_cnt, true);
                grid.sync();
            }
            td_fwd_body<3, NW>(S, d, a.vars, A, st, in, out, ld, n, inv_cnt, a.batch_stats != 0);
        } else {
            td_scale_fwd_body(sh, a.tables + op.sidx * NF_MAX_ROWS * 2, op.o.token != NF_TOKEN_GAIN4, 1, in, (const float4*)a.y, out, ld,
                              a.rows, a.default_row, n);
        }
        __syncthreads();        // the next op reads what this CTA just wrote
    }
    td_nll_body(sh, (const float4*)slot(G - 1), ld, a.consts, (float4*)gA, nll, sdz, n, inv_n);
    __syncthreads();
    // ---- backward
    for (int i = G - 1; i >= 0; --i) {
        const TdOp& op = prog->ops[i];
        const float4* zin = (const float4*)(i == 0 ? a.x : slot(i - 1));
        const float4* zout = (const float4*)slot(i);
        if (op.o.kind == NF_TOP_COUPLING) {
            const TdCoupling d = td_desc(op.o, a.batch_stats, a.bn_eps);
            const float* A = a.Amat + op.cidx * 16;
            const double* st = a.stats + op.cidx * 16;
            double* cgr = a.cgrads + (long long)op.cidx * NF_G_COUPLING_DOUBLES;
            td_b1_body<NW>(S, d, a.vars, A, st, zin, (const float4*)gA, (float4*)gzp, (float4*)scratch, n, inv_n, inv_cnt, cgr);
            if (a.batch_stats) grid.sync(); else __syncthreads();      // BatchNorm-2 backward sums over the whole batch
            td_b2_body<NW>(S, d, a.vars, A, st, zin, (float4*)scratch, n, inv_cnt, cgr, true);
            if (a.batch_stats) grid.sync(); else __syncthreads();      // BatchNorm-1 backward sums
            td_b3_body<NW>(S, d, a.vars, A, st, zin, (const float4*)scratch, (const float4*)gzp, (float4*)gB, n, inv_cnt, cgr, true);
        } else {
            td_scale_bwd_body(sh, a.tables + op.sidx * NF_MAX_ROWS * 2, op.o.token != NF_TOKEN_GAIN4, 1, zout, (const float4*)a.y,
                              (const float4*)gA, (float4*)gB, a.rows, a.default_row, n, inv_n, a.sgrads + (long long)op.sidx * NF_MAX_ROWS * 2);
        }
        __syncthreads();
        float* tmp = gA; gA = gB; gB = tmp;
    }
}

// ---------------------------------------------------------------------------------------------- small-batch chain, batch statistics
// nf_chain_batch_stats (NoiseFlow / NoiseFlowWrapper with is_training == True) for batches that fit one co-resident CTA per
// patch: the whole chain, both BatchNorm probes of every coupling included, as ONE cooperative kernel built from the pass
// bodies above -- instead of three launches, two device-to-host copies and two stream synchronisations per coupling.  This is
// the reference sampling script's call pattern (sample_noise_flow.py:44,71: batch_size = 1, one sess.run per patch).
//   direction 0 (data -> latent): per op F1 | F2 | F3 exactly as the trainer's forward, then prior / NLL.
//   direction 1 (latent -> data): ops arrive reversed; a coupling's net sees its INPUT's first two channels (they pass
//     through unchanged), so F1 / F2 run unchanged on the un-mixed patch, then x1 = (y1 - shift) exp(-ls) and the 1x1 conv /
//     permutation is applied last with the inverse matrix.
// last stage of a coupling in the sampling direction: y1 -> (y1 - shift) * exp(-ls), then the mix with the inverse matrix
template <int NW>
__device__ __forceinline__ void td_sample3_body(TdSmem& S, const TdCoupling& d, const float* __restrict__ vars, const float* Ainv,
                                                int has_mix, const double* stats, const float4* zin, float4* zout, float* ld, long long n,
                                                double inv_cnt, bool resident) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // resident (one patch per CTA): weights, the (un-mixed) patch and its conv-1 output are still in shared memory from F2
    td_load_params(S, d, vars, nullptr, stats, inv_cnt, 2, resident);
    __shared__ float Ai[16];
    if (threadIdx.x < 16) Ai[threadIdx.x] = has_mix ? __ldcg(Ainv + threadIdx.x) : ((threadIdx.x >> 2) == (threadIdx.x & 3) ? 1.f : 0.f);
    __syncthreads();
    float e3[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) e3[o] = expf(3.f * S.P.logs[o]);
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        if (!resident) td_load_mixed<NW>(S, zin + p * NF_PIXELS, warp, lane);      // d.has_mix == 0 here: a plain copy
        for (int k = threadIdx.x; k < 34 * 34; k += blockDim.x)
            if (on_ring(k)) S.h2[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        if (!resident) td_conv1_rows<NW>(S, warp, lane);
        for (int r = warp; r < 32; r += NW) {
            float c1hat[4], h1[4], c2hat[4];
            td_net_to_c2hat(S, r, lane, c1hat, h1, c2hat);
            S.h2[(r + 1) * 34 + lane + 1] = make_float4(fmaxf(c2hat[0], 0.f), fmaxf(c2hat[1], 0.f), fmaxf(c2hat[2], 0.f), fmaxf(c2hat[3], 0.f));
        }
        __syncthreads();
        float pre_rows[32 / NW][4];
        td_conv3_rows<NW>(S, warp, lane, pre_rows);
        float lsum = 0.f;
#pragma unroll
        for (int k = 0; k < 32 / NW; ++k) {
            const int r = warp + k * NW;
            const float (&pre)[4] = pre_rows[k];
            const float sh0 = pre[0] * e3[0], sh1 = pre[1] * e3[1];
            const float ls0 = S.P.scale * tanhf(pre[2] * e3[2]), ls1 = S.P.scale * tanhf(pre[3] * e3[3]);
            const float4 zp = S.zp[r * 32 + lane];
            float4 v = make_float4(zp.x, zp.y, (zp.z - sh0) * expf(-ls0), (zp.w - sh1) * expf(-ls1));     // layers.py:296-301
            if (has_mix) v = mix_fwd(v, Ai);
            zout[p * NF_PIXELS + r * 32 + lane] = v;
            lsum -= ls0 + ls1;
        }
        if (ld) {
            lsum = tw_sum(lsum);
            if (lane == 0) S.red[warp] = lsum;
            __syncthreads();
            if (threadIdx.x == 0) {
                float tot = 0.f;
                for (int w = 0; w < NW; ++w) tot += S.red[w];
                ld[p] += tot;
            }
        }
        __syncthreads();
    }
}

// scale layer in the sampling direction: z * scale (AffineCouplingSdnEx5._forward etc.)
__device__ __forceinline__ void td_scale_sample_body(const float* table, int is_sdn, const float4* zin, const float4* __restrict__ y,
                                                     float4* zout, const int* __restrict__ rows, int default_row, long long n) {
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        int row = rows ? rows[p] : default_row;
        row = min(max(row, 0), NF_MAX_ROWS - 1);
        const float a = __ldcg(table + row * 2), b = __ldcg(table + row * 2 + 1);
        for (int k = threadIdx.x; k < NF_PIXELS; k += blockDim.x) {
            const long long idx = p * NF_PIXELS + k;
            const float4 z = zin[idx];
            if (is_sdn) {
                const float4 yv = y[idx];
                zout[idx] = make_float4(z.x * sqrtf(fmaf(a, yv.x, b)), z.y * sqrtf(fmaf(a, yv.y, b)), z.z * sqrtf(fmaf(a, yv.z, b)),
                                        z.w * sqrtf(fmaf(a, yv.w, b)));
            } else {
                zout[idx] = make_float4(z.x * a, z.y * a, z.z * a, z.w * a);
            }
        }
    }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, 2)
td_bs_chain_kernel(const BsArgs a) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TdSmem& S = *reinterpret_cast<TdSmem*>(smem_raw);
    __shared__ float sh[16];
    const long long n = a.n;
    const double inv_cnt = 1.0 / ((double)n * NF_PIXELS);
    const float4* cur = (const float4*)a.in;
    float4* out = (float4*)a.out;
    // one patch per CTA: the patch, the weights and the conv-1 output stay in shared memory between the passes of a coupling;
    // more patches than co-resident CTAs: every pass walks its patches grid-stride and re-loads (still one launch, no host)
    const bool resident = n <= (long long)gridDim.x;
    if (a.direction == 1) {     // z = eps * temp (noise_flow_model.py:499-504); eps given or Philox4x32-10 as nf_sample
        for (long long p = blockIdx.x; p < n; p += gridDim.x)
            for (int k = threadIdx.x; k < NF_PIXELS; k += blockDim.x) {
                float4 v = cur ? cur[p * NF_PIXELS + k] : philox_normal4(a.seed, a.offset, a.patch_base + (unsigned long long)p, (unsigned int)k);
                v.x *= a.temp; v.y *= a.temp; v.z *= a.temp; v.w *= a.temp;
                out[p * NF_PIXELS + k] = v;
            }
        __syncthreads();
        cur = out;
    }
    for (int i = 0; i < a.n_ops; ++i) {
        const BsOp& op = a.ops[i];
        if (op.kind == 0) {
            TdCoupling d = op.d;
            const float* A = a.Amat + op.cidx * 16;
            double* st = a.stats + op.cidx * 16;
            const int has_mix = d.has_mix;
            if (a.direction == 1) d.has_mix = 0;          // the net sees the un-mixed input; the mix comes last
            td_fwd_body<1, NW>(S, d, a.vars, A, st, cur, out, a.ld, n, inv_cnt);
            grid.sync();
            td_fwd_body<2, NW>(S, d, a.vars, A, st, cur, out, a.ld, n, inv_cnt, resident);
            grid.sync();
            if (a.direction == 0) td_fwd_body<3, NW>(S, d, a.vars, A, st, cur, out, a.ld, n, inv_cnt, resident);
            else td_sample3_body<NW>(S, d, a.vars, A, has_mix, st, cur, out, a.ld, n, inv_cnt, resident);
        } else if (a.direction == 0) {
            td_scale_fwd_body(sh, a.tables + op.sidx * NF_MAX_ROWS * 2, op.is_sdn, op.full_sum, cur, (const float4*)a.y, out, a.ld, a.rows,
                              a.default_row, n);
        } else {
            td_scale_sample_body(a.tables + op.sidx * NF_MAX_ROWS * 2, op.is_sdn, cur, (const float4*)a.y, out, a.rows, a.default_row, n);
        }
        __syncthreads();
        cur = out;
    }
    if (a.direction == 0 && (a.nll || a.sdz || a.logdet)) {
        if (a.logdet && a.logdet != a.ld)
            for (long long p = blockIdx.x; p < n; p += gridDim.x)
                if (threadIdx.x == 0) a.logdet[p] = a.ld[p] + (float)__ldcg(a.consts);
        td_nll_body(sh, cur, a.ld, a.consts, nullptr, a.nll, a.sdz, n, 0.f);
        if (a.logdet && a.logdet == a.ld) {
            __syncthreads();
            for (long long p = blockIdx.x; p < n; p += gridDim.x)
                if (threadIdx.x == 0) a.logdet[p] = a.ld[p] + (float)__ldcg(a.consts);
        }
    }
}

// Measurement aid for the roofline of the fused step (bench.py): `reps` grid-wide barriers of a cooperative grid with the
// shape of td_step_kernel<8> (256 threads, TdSmem of dynamic shared memory) and nothing else.
__global__ void __launch_bounds__(256, 2) td_barrier_probe_kernel(int reps, int* sink) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    int v = 0;
    for (int r = 0; r < reps; ++r) { grid.sync(); v += r; }
    if (sink && v == -1) *sink = v;
}
cudaError_t launch_barrier_probe(int n_ctas, int reps, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(td_barrier_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TdSmem));
    if (e != cudaSuccess) return e;
    int* sink = nullptr;
    void* kargs[] = {(void*)&reps, (void*)&sink};
    return cudaLaunchCooperativeKernel((const void*)td_barrier_probe_kernel, dim3((unsigned)n_ctas), dim3(256), kargs, sizeof(TdSmem), s);
}

int bs_small_capacity(int sm_count) {
    static int per_sm_dev[NF_MAX_DEVICES] = {};   // per device; 0 = not queried yet, stored as occupancy + 1
    int& slot = per_sm_dev[device_slot()];
    int per_sm = slot - 1;
    if (per_sm < 0) {
        if (cudaFuncSetAttribute(td_bs_chain_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TdSmem)) != cudaSuccess) return 0;
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, td_bs_chain_kernel<8>, 256, sizeof(TdSmem)) != cudaSuccess) return 0;
        per_sm = v;
        slot = v + 1;
    }
    return per_sm * sm_count;
}

cudaError_t launch_bs_small(const BsArgs& a, int capacity, cudaStream_t s) {
    void* kargs[] = {(void*)&a};
    const long long grid = a.n < (long long)capacity ? a.n : (long long)capacity;      // beyond: grid-stride over the patches
    return cudaLaunchCooperativeKernel((const void*)td_bs_chain_kernel<8>, dim3((unsigned)grid), dim3(256), kargs, sizeof(TdSmem), s);
}

}  // namespace nf

// =====================================================================================================
// host side
// =====================================================================================================
struct nf_trainer {
    nf::TdProgram prog = {};
    int n_cp = 0, n_sc = 0;
    int64_t n_vars = 0, max_batch = 0;
    float bn_eps = 1e-4f;
    int sm_count = 0;
    // device memory
    nf::TdProgram* d_prog = nullptr;
    float* d_vars = nullptr;
    unsigned char* d_trainable = nullptr;
    double *d_am = nullptr, *d_av = nullptr;
    long long* d_step = nullptr;
    float* d_ws = nullptr;        // (n_ops + 4) activation / gradient slots + ld, nll, sdz
    double* d_dbl = nullptr;      // [stats 16 n_cp][cgrads 320 n_cp][sgrads 64 n_sc][consts 4]
    float* d_A = nullptr;         // derived 1x1 matrices, 16 per coupling
    float* d_tables = nullptr;    // scale tables, [n_sc][32][2]
    int64_t dbl_len = 0;
    // CUDA-graph replay of the loss+gradient launch sequence: inputs are staged into trainer-owned buffers so the
    // captured pointers never change; the graph is re-captured only when (n, row, mode, reduce buffer) change
    int cta_warps = 0;            // 0 = automatic (= 8); 8 / 16 = forced (nf_trainer_set_cta_warps)
    int fused = 1;                // 1 = one cooperative kernel per loss+gradient when the batch is co-resident (nf_trainer_set_fused)
    int coop_cap[2] = {-1, -1};   // co-resident CTAs of td_step_kernel<8>, <16> (-1 = not queried yet)
    int last_launches = 0, last_mode = -1;   // kernels enqueued by the last loss+gradient, and its batch_stats flag
    int use_graph = 1;
    float *d_xs = nullptr, *d_ys = nullptr;
    int32_t* d_rows = nullptr;
    cudaStream_t cap = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    cudaGraphExec_t exec = nullptr;
    struct Key { int64_t n; int32_t row; int batch_stats, has_y, has_rows; double* red; } key = {};
};

namespace {
#define TD_CUDA(call)                                                                                              \
    do {                                                                                                           \
        cudaError_t e__ = (call);                                                                                  \
        if (e__ != cudaSuccess) return nf::set_error(NF_ERR_CUDA, #call, cudaGetErrorString(e__));                 \
    } while (0)

int td_smem_attr() {
    static bool done_dev[NF_MAX_DEVICES] = {};   // per device
    bool& done = done_dev[nf::device_slot()];
    if (done) return NF_OK;
    const int bytes = (int)sizeof(nf::TdSmem);
#define TD_ATTR(K) TD_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
    TD_ATTR((nf::td_fwd_kernel<1, 8>)); TD_ATTR((nf::td_fwd_kernel<2, 8>)); TD_ATTR((nf::td_fwd_kernel<3, 8>));
    TD_ATTR((nf::td_fwd_kernel<1, 16>)); TD_ATTR((nf::td_fwd_kernel<2, 16>)); TD_ATTR((nf::td_fwd_kernel<3, 16>));
    TD_ATTR(nf::td_b1_kernel<8>); TD_ATTR(nf::td_b2_kernel<8>); TD_ATTR(nf::td_b3_kernel<8>);
    TD_ATTR(nf::td_b1_kernel<16>); TD_ATTR(nf::td_b2_kernel<16>); TD_ATTR(nf::td_b3_kernel<16>);
    TD_ATTR(nf::td_step_kernel<8>); TD_ATTR(nf::td_step_kernel<16>);
#undef TD_ATTR
    done = true;
    return NF_OK;
}

nf::TdCoupling coupling_desc(const nf_trainer* t, const nf_train_op& o, int batch_stats) {
    nf::TdCoupling d = {};
    d.off_w1 = o.off_w1; d.off_b1 = o.off_b1; d.off_w2 = o.off_w2; d.off_b2 = o.off_b2;
    d.off_w3 = o.off_w3; d.off_b3 = o.off_b3; d.off_logs = o.off_logs; d.off_scale = o.off_scale;
    d.off_bn[0] = o.off_bn1_mean; d.off_bn[1] = o.off_bn1_var; d.off_bn[2] = o.off_bn2_mean; d.off_bn[3] = o.off_bn2_var;
    d.has_mix = o.mix_kind != 0;
    d.batch_stats = batch_stats ? 1 : 0;
    d.bn_eps = t->bn_eps;
    return d;
}

bool off_ok(int32_t off, int64_t len, int64_t n_vars) { return off >= 0 && (int64_t)off + len <= n_vars; }
}  // namespace

extern "C" {

int nf_trainer_create(const nf_train_op* ops, int n_ops, const float* vars_host, const uint8_t* trainable, int64_t n_vars,
                      int64_t max_batch, const int32_t* tri_lower, const int32_t* tri_upper, float bn_eps, nf_trainer** out) {
    if (!ops || !vars_host || !trainable || !tri_lower || !tri_upper || !out) return nf::set_error(NF_ERR_INVALID, "nf_trainer_create", "null argument");
    if (n_ops < 1 || n_ops > TD_MAX_OPS) return nf::set_error(NF_ERR_UNSUPPORTED, "nf_trainer_create", "number of ops outside 1..40");
    if (n_vars < 1 || max_batch < 1) return nf::set_error(NF_ERR_INVALID, "nf_trainer_create", "n_vars and max_batch must be positive");
    nf_trainer* t = new (std::nothrow) nf_trainer();
    if (!t) return nf::set_error(NF_ERR_INVALID, "nf_trainer_create", "out of host memory");
    t->n_vars = n_vars; t->max_batch = max_batch; t->bn_eps = bn_eps;
    t->prog.n_ops = n_ops;
    for (int k = 0; k < 6; ++k) {
        if (tri_lower[k] < 0 || tri_lower[k] > 15 || tri_upper[k] < 0 || tri_upper[k] > 15) { delete t; return nf::set_error(NF_ERR_INVALID, "nf_trainer_create", "triangle position outside 0..15"); }
        t->prog.tri_lo[k] = tri_lower[k]; t->prog.tri_up[k] = tri_upper[k];
    }
    for (int i = 0; i < n_ops; ++i) {
        const nf_train_op& o = ops[i];
        nf::TdOp& d = t->prog.ops[i];
        d.o = o; d.cidx = -1; d.sidx = -1;
        bool ok = true;
        if (o.kind == NF_TOP_COUPLING) {
            d.cidx = t->n_cp++;
            ok = off_ok(o.off_w1, 72, n_vars) && off_ok(o.off_b1, 4, n_vars) && off_ok(o.off_w2, 16, n_vars) && off_ok(o.off_b2, 4, n_vars) &&
                 off_ok(o.off_w3, 180, n_vars) && off_ok(o.off_b3, 4, n_vars) && off_ok(o.off_logs, 4, n_vars) && off_ok(o.off_scale, 1, n_vars) &&
                 off_ok(o.off_bn1_mean, 4, n_vars) && off_ok(o.off_bn1_var, 4, n_vars) && off_ok(o.off_bn2_mean, 4, n_vars) && off_ok(o.off_bn2_var, 4, n_vars);
            if (o.mix_kind == 1)
                ok = ok && off_ok(o.off_P, 16, n_vars) && off_ok(o.off_L, 6, n_vars) && off_ok(o.off_U, 6, n_vars) && off_ok(o.off_logS, 4, n_vars) && off_ok(o.off_signS, 4, n_vars);
            else if (o.mix_kind == 2) {
                bool seen[4] = {false, false, false, false};
                for (int k = 0; k < 4; ++k) { if (o.perm[k] < 0 || o.perm[k] > 3 || seen[o.perm[k]]) ok = false; else seen[o.perm[k]] = true; }
            } else if (o.mix_kind != 0) ok = false;
        } else if (o.kind == NF_TOP_SCALE) {
            d.sidx = t->n_sc++;
            if (o.token == NF_TOKEN_GAIN4) ok = off_ok(o.off_gain_val, 1, n_vars);
            else if (o.token == NF_TOKEN_SDN4 || o.token == NF_TOKEN_SDN5 || o.token == NF_TOKEN_SDN6) {
                ok = off_ok(o.off_beta1, 1, n_vars) && off_ok(o.off_beta2, 1, n_vars) && off_ok(o.off_gain_params, 5, n_vars);
                if (o.token == NF_TOKEN_SDN5) ok = ok && off_ok(o.off_cam_params, 15, n_vars);
                if (o.token == NF_TOKEN_SDN6) ok = ok && off_ok(o.off_cam_params, 5, n_vars);
            } else ok = false;
        } else ok = false;
        if (!ok) { delete t; return nf::set_error(NF_ERR_INVALID, "nf_trainer_create", "op has an unknown kind / token or an offset outside the variable array"); }
    }
    int rc = nf_device_info(&t->sm_count, nullptr, nullptr, nullptr);
    if (rc) { delete t; return rc; }
    t->dbl_len = 16 * (int64_t)t->n_cp + NF_G_COUPLING_DOUBLES * (int64_t)t->n_cp + 2 * NF_MAX_ROWS * (int64_t)t->n_sc + 4;
    const size_t ws_floats = (size_t)(n_ops + 4) * (size_t)max_batch * NF_DIMS + 3 * (size_t)max_batch;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes ? bytes : 16); };
    alloc((void**)&t->d_prog, sizeof(nf::TdProgram));
    alloc((void**)&t->d_vars, (size_t)n_vars * sizeof(float));
    alloc((void**)&t->d_trainable, (size_t)n_vars);
    alloc((void**)&t->d_am, (size_t)n_vars * sizeof(double));
    alloc((void**)&t->d_av, (size_t)n_vars * sizeof(double));
    alloc((void**)&t->d_step, sizeof(long long));
    alloc((void**)&t->d_ws, ws_floats * sizeof(float));
    alloc((void**)&t->d_dbl, (size_t)t->dbl_len * sizeof(double));
    alloc((void**)&t->d_A, (size_t)(t->n_cp ? t->n_cp : 1) * 16 * sizeof(float));
    alloc((void**)&t->d_tables, (size_t)(t->n_sc ? t->n_sc : 1) * NF_MAX_ROWS * 2 * sizeof(float));
    alloc((void**)&t->d_xs, (size_t)max_batch * NF_DIMS * sizeof(float));
    alloc((void**)&t->d_ys, (size_t)max_batch * NF_DIMS * sizeof(float));
    alloc((void**)&t->d_rows, (size_t)max_batch * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&t->cap, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_in, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_out, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMemcpy(t->d_prog, &t->prog, sizeof(nf::TdProgram), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->d_vars, vars_host, (size_t)n_vars * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->d_trainable, trainable, (size_t)n_vars, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(t->d_am, 0, (size_t)n_vars * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(t->d_av, 0, (size_t)n_vars * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(t->d_step, 0, sizeof(long long));
    if (e != cudaSuccess) {
        nf_trainer_destroy(t);
        return nf::set_error(NF_ERR_CUDA, "nf_trainer_create", cudaGetErrorString(e));
    }
    *out = t;
    return NF_OK;
}

int nf_trainer_destroy(nf_trainer* t) {
    if (!t) return NF_OK;
    cudaFree(t->d_prog); cudaFree(t->d_vars); cudaFree(t->d_trainable); cudaFree(t->d_am); cudaFree(t->d_av);
    cudaFree(t->d_step); cudaFree(t->d_ws); cudaFree(t->d_dbl); cudaFree(t->d_A); cudaFree(t->d_tables);
    cudaFree(t->d_xs); cudaFree(t->d_ys); cudaFree(t->d_rows);
    if (t->exec) cudaGraphExecDestroy(t->exec);
    if (t->ev_in) cudaEventDestroy(t->ev_in);
    if (t->ev_out) cudaEventDestroy(t->ev_out);
    if (t->cap) cudaStreamDestroy(t->cap);
    delete t;
    return NF_OK;
}

int nf_trainer_reduce_len(const nf_trainer* t, int64_t* n_doubles) {
    if (!t || !n_doubles) return nf::set_error(NF_ERR_INVALID, "nf_trainer_reduce_len", "null argument");
    *n_doubles = t->n_vars + 3 + 16 * (int64_t)t->n_cp;
    return NF_OK;
}

int nf_probe_grid_barrier(int n_ctas, int reps, float* us_per_barrier, void* stream) {
    if (n_ctas < 1 || reps < 1 || !us_per_barrier) return nf::set_error(NF_ERR_INVALID, "nf_probe_grid_barrier", "bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    TD_CUDA(cudaEventCreate(&e0));
    TD_CUDA(cudaEventCreate(&e1));
    cudaError_t e = nf::launch_barrier_probe(n_ctas, 8, s);          // warm-up
    float ms_a = 0.f, ms_b = 0.f;
    if (e == cudaSuccess) e = cudaEventRecord(e0, s);
    if (e == cudaSuccess) e = nf::launch_barrier_probe(n_ctas, reps, s);
    if (e == cudaSuccess) e = cudaEventRecord(e1, s);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms_a, e0, e1);
    if (e == cudaSuccess) e = cudaEventRecord(e0, s);
    if (e == cudaSuccess) e = nf::launch_barrier_probe(n_ctas, 2 * reps, s);
    if (e == cudaSuccess) e = cudaEventRecord(e1, s);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms_b, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) return nf::set_error(NF_ERR_CUDA, "nf_probe_grid_barrier", cudaGetErrorString(e));
    *us_per_barrier = (ms_b - ms_a) * 1e3f / (float)reps;        // the difference removes the launch cost
    return NF_OK;
}

int nf_trainer_barriers_per_step(const nf_trainer* t, int batch_stats, int* n_barriers) {
    if (!t || !n_barriers) return nf::set_error(NF_ERR_INVALID, "nf_trainer_barriers_per_step", "null argument");
    // fused step: one after prep, two per coupling and direction around the BatchNorm batch sums
    *n_barriers = 1 + (batch_stats ? 4 * t->n_cp : 0);
    return NF_OK;
}

int nf_trainer_launches_per_step(const nf_trainer* t, int batch_stats, int* n_launches) {
    if (!t || !n_launches) return nf::set_error(NF_ERR_INVALID, "nf_trainer_launches_per_step", "null argument");
    // prep + per coupling (F1, F2 with batch statistics) F3, B1, B2, B3 + per scale layer fwd, bwd + nll + reduce + chain + apply
    *n_launches = 1 + t->n_cp * ((batch_stats ? 2 : 0) + 4) + t->n_sc * 2 + 4;
    // what the last evaluation in this mode really enqueued (the fused path: step + reduce + chain), + apply
    if (t->last_launches > 0 && t->last_mode == (batch_stats ? 1 : 0)) *n_launches = t->last_launches + 1;
    return NF_OK;
}

// the launch sequence of one loss+gradient evaluation (no synchronisation; capturable)
static int td_enqueue(nf_trainer* t, const float* x, const float* y, const int32_t* rows, int32_t default_row, int64_t n,
                      int batch_stats, double* red, cudaStream_t s) {
    const int G = t->prog.n_ops;
    const int64_t S = n * NF_DIMS;
    auto slot = [&](int k) { return t->d_ws + (int64_t)k * S; };
    float *gA = slot(G), *gB = slot(G + 1), *scratch = slot(G + 2), *gzp = slot(G + 3);
    float *d_ld = slot(G + 4), *d_nll = d_ld + n, *d_sdz = d_nll + n;
    double* d_stats = t->d_dbl;
    double* d_cg = d_stats + 16 * (int64_t)t->n_cp;
    double* d_sg = d_cg + NF_G_COUPLING_DOUBLES * (int64_t)t->n_cp;
    double* d_consts = d_sg + 2 * NF_MAX_ROWS * (int64_t)t->n_sc;
    const double inv_cnt = 1.0 / ((double)n * NF_PIXELS);
    const float inv_n = 1.f / (float)n;
    const unsigned grid = (unsigned)n;
    const size_t smem = sizeof(nf::TdSmem);
    const bool wide_cta = t->cta_warps == 16;      // automatic = 8: four rows per warp amortise every weight load (0.465 vs 0.489 ms at 138)
    for (int i = 0; i < G; ++i)
        if (t->prog.ops[i].o.kind == NF_TOP_SCALE && t->prog.ops[i].o.token != NF_TOKEN_GAIN4 && !y)
            return nf::set_error(NF_ERR_INVALID, "nf_trainer_loss_and_grad", "clean patch y is required by an sdn layer");

    TD_CUDA(cudaMemsetAsync(t->d_dbl, 0, (size_t)t->dbl_len * sizeof(double), s));
    TD_CUDA(cudaMemsetAsync(red, 0, (size_t)(t->n_vars + 3 + 16 * t->n_cp) * sizeof(double), s));
    TD_CUDA(cudaMemsetAsync(d_ld, 0, (size_t)n * sizeof(float), s));
    t->last_mode = batch_stats ? 1 : 0;
    if (t->fused) {
        // one cooperative kernel when every patch-CTA can be resident at once (296 at 8 warps, 148 at 16 on a B200)
        int& cap = t->coop_cap[wide_cta ? 1 : 0];
        if (cap < 0) {
            int per_sm = 0;
            cudaError_t e = wide_cta ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nf::td_step_kernel<16>, 512, smem)
                                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nf::td_step_kernel<8>, 256, smem);
            if (e != cudaSuccess) return nf::set_error(NF_ERR_CUDA, "cudaOccupancyMaxActiveBlocksPerMultiprocessor", cudaGetErrorString(e));
            cap = per_sm * t->sm_count;
        }
        if (n <= (int64_t)cap) {
            nf::TdStepArgs a = {};
            a.prog = t->d_prog; a.vars = t->d_vars; a.Amat = t->d_A; a.tables = t->d_tables;
            a.stats = d_stats; a.cgrads = d_cg; a.sgrads = d_sg; a.consts = d_consts;
            a.ws = t->d_ws; a.stride = S; a.x = x; a.y = y; a.rows = rows; a.default_row = default_row;
            a.n = n; a.batch_stats = batch_stats ? 1 : 0; a.bn_eps = t->bn_eps;
            void* kargs[] = {&a};
            TD_CUDA(cudaLaunchCooperativeKernel(wide_cta ? (const void*)nf::td_step_kernel<16> : (const void*)nf::td_step_kernel<8>,
                                                dim3(grid), dim3(wide_cta ? 512 : 256), kargs, smem, s));
            cudaError_t e = nf::launch_reduce(d_nll, d_sdz, n, red + t->n_vars, s);
            if (e != cudaSuccess) return nf::set_error(NF_ERR_CUDA, "reduce launch", cudaGetErrorString(e));
            nf::td_chain_kernel<<<1, 512, 0, s>>>(t->d_prog, t->d_vars, d_cg, d_sg, d_stats, inv_cnt, batch_stats ? 1 : 0, t->n_vars, red);
            TD_CUDA(cudaGetLastError());
            t->last_launches = 3;
            return NF_OK;
        }
    }
    t->last_launches = 1 + t->n_cp * ((batch_stats ? 2 : 0) + 4) + t->n_sc * 2 + 3;
    nf::td_prep_kernel<<<1, 256, 0, s>>>(t->d_prog, t->d_vars, t->d_A, t->d_tables, d_consts);
    // ---- forward, keeping every op's input
    for (int i = 0; i < G; ++i) {
        const nf::TdOp& op = t->prog.ops[i];
        const float4* in = (const float4*)(i == 0 ? x : slot(i - 1));
        float4* out = (float4*)slot(i);
        if (op.o.kind == NF_TOP_COUPLING) {
            const nf::TdCoupling d = coupling_desc(t, op.o, batch_stats);
            const float* A = t->d_A + op.cidx * 16;
            double* st = d_stats + op.cidx * 16;
            if (wide_cta) {
                if (batch_stats) {
                    nf::td_fwd_kernel<1, 16><<<grid, 512, smem, s>>>(d, t->d_vars, A, st, in, out, d_ld, n, inv_cnt);
                    nf::td_fwd_kernel<2, 16><<<grid, 512, smem, s>>>(d, t->d_vars, A, st, in, out, d_ld, n, inv_cnt);
                }
                nf::td_fwd_kernel<3, 16><<<grid, 512, smem, s>>>(d, t->d_vars, A, st, in, out, d_ld, n, inv_cnt);
            } else {
                if (batch_stats) {
                    nf::td_fwd_kernel<1, 8><<<grid, 256, smem, s>>>(d, t->d_vars, A, st, in, out, d_ld, n, inv_cnt);
                    nf::td_fwd_kernel<2, 8><<<grid, 256, smem, s>>>(d, t->d_vars, A, st, in, out, d_ld, n, inv_cnt);
                }
                nf::td_fwd_kernel<3, 8><<<grid, 256, smem, s>>>(d, t->d_vars, A, st, in, out, d_ld, n, inv_cnt);
            }
        } else {
            const int is_sdn = op.o.token != NF_TOKEN_GAIN4;
            nf::td_scale_fwd_kernel<<<grid, TD_THREADS, 0, s>>>(t->d_tables + op.sidx * NF_MAX_ROWS * 2, is_sdn, 1, in, (const float4*)y, out,
                                                                d_ld, rows, default_row, n);
        }
    }
    nf::td_nll_kernel<<<grid, TD_THREADS, 0, s>>>((const float4*)slot(G - 1), d_ld, d_consts, (float4*)gA, d_nll, d_sdz, n, inv_n);
    cudaError_t e = nf::launch_reduce(d_nll, d_sdz, n, red + t->n_vars, s);
    if (e != cudaSuccess) return nf::set_error(NF_ERR_CUDA, "reduce launch", cudaGetErrorString(e));
    // ---- backward
    for (int i = G - 1; i >= 0; --i) {
        const nf::TdOp& op = t->prog.ops[i];
        const float4* zin = (const float4*)(i == 0 ? x : slot(i - 1));
        const float4* zout = (const float4*)slot(i);
        if (op.o.kind == NF_TOP_COUPLING) {
            const nf::TdCoupling d = coupling_desc(t, op.o, batch_stats);
            const float* A = t->d_A + op.cidx * 16;
            const double* st = d_stats + op.cidx * 16;
            double* cg = d_cg + (int64_t)op.cidx * NF_G_COUPLING_DOUBLES;
            if (wide_cta) {
                nf::td_b1_kernel<16><<<grid, 512, smem, s>>>(d, t->d_vars, A, st, zin, (const float4*)gA, (float4*)gzp, (float4*)scratch, n, inv_n, inv_cnt, cg);
                nf::td_b2_kernel<16><<<grid, 512, smem, s>>>(d, t->d_vars, A, st, zin, (float4*)scratch, n, inv_cnt, cg);
                nf::td_b3_kernel<16><<<grid, 512, smem, s>>>(d, t->d_vars, A, st, zin, (const float4*)scratch, (const float4*)gzp, (float4*)gB, n, inv_cnt, cg);
            } else {
                nf::td_b1_kernel<8><<<grid, 256, smem, s>>>(d, t->d_vars, A, st, zin, (const float4*)gA, (float4*)gzp, (float4*)scratch, n, inv_n, inv_cnt, cg);
                nf::td_b2_kernel<8><<<grid, 256, smem, s>>>(d, t->d_vars, A, st, zin, (float4*)scratch, n, inv_cnt, cg);
                nf::td_b3_kernel<8><<<grid, 256, smem, s>>>(d, t->d_vars, A, st, zin, (const float4*)scratch, (const float4*)gzp, (float4*)gB, n, inv_cnt, cg);
            }
        } else {
            const int is_sdn = op.o.token != NF_TOKEN_GAIN4;
            nf::td_scale_bwd_kernel<<<grid, TD_THREADS, 0, s>>>(t->d_tables + op.sidx * NF_MAX_ROWS * 2, is_sdn, 1, zout, (const float4*)y,
                                                                (const float4*)gA, (float4*)gB, rows, default_row, n, inv_n,
                                                                d_sg + (int64_t)op.sidx * NF_MAX_ROWS * 2);
        }
        float* tmp = gA; gA = gB; gB = tmp;
    }
    nf::td_chain_kernel<<<1, 512, 0, s>>>(t->d_prog, t->d_vars, d_cg, d_sg, d_stats, inv_cnt, batch_stats ? 1 : 0, t->n_vars, red);
    TD_CUDA(cudaGetLastError());
    return NF_OK;
}

int nf_trainer_set_cta_warps(nf_trainer* t, int warps) {
    if (!t || (warps != 0 && warps != 8 && warps != 16)) return nf::set_error(NF_ERR_INVALID, "nf_trainer_set_cta_warps", "warps must be 0 (automatic), 8 or 16");
    t->cta_warps = warps;
    if (t->exec) { cudaGraphExecDestroy(t->exec); t->exec = nullptr; }   // the captured launch shapes are stale
    return NF_OK;
}

int nf_trainer_set_fused(nf_trainer* t, int enable) {
    if (!t) return nf::set_error(NF_ERR_INVALID, "nf_trainer_set_fused", "null trainer");
    t->fused = enable ? 1 : 0;
    if (t->exec) { cudaGraphExecDestroy(t->exec); t->exec = nullptr; }
    return NF_OK;
}

int nf_trainer_set_graph(nf_trainer* t, int enable) {
    if (!t) return nf::set_error(NF_ERR_INVALID, "nf_trainer_set_graph", "null trainer");
    t->use_graph = enable ? 1 : 0;
    return NF_OK;
}

int nf_trainer_loss_and_grad(nf_trainer* t, const float* x, const float* y, const int32_t* rows, int32_t default_row, int64_t n,
                             int batch_stats, double* red, void* stream_) {
    if (!t || !x || !red) return nf::set_error(NF_ERR_INVALID, "nf_trainer_loss_and_grad", "trainer, x and reduce_buf are required");
    if (n < 1 || n > t->max_batch) return nf::set_error(NF_ERR_INVALID, "nf_trainer_loss_and_grad", "batch size outside 1..max_batch");
    if (default_row < 0 || default_row >= 25) return nf::set_error(NF_ERR_INVALID, "nf_trainer_loss_and_grad", "default_row outside the standard (camera, ISO) grid 0..24");
    int rc = td_smem_attr();
    if (rc) return rc;
    cudaStream_t user = (cudaStream_t)stream_;
    if (!t->use_graph) return td_enqueue(t, x, y, rows, default_row, n, batch_stats, red, user);
    // stage the inputs (device -> device, stream-ordered) so that the captured graph sees constant pointers
    TD_CUDA(cudaMemcpyAsync(t->d_xs, x, (size_t)n * NF_DIMS * sizeof(float), cudaMemcpyDeviceToDevice, user));
    if (y) TD_CUDA(cudaMemcpyAsync(t->d_ys, y, (size_t)n * NF_DIMS * sizeof(float), cudaMemcpyDeviceToDevice, user));
    if (rows) TD_CUDA(cudaMemcpyAsync(t->d_rows, rows, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, user));
    const nf_trainer::Key key = {n, default_row, batch_stats ? 1 : 0, y ? 1 : 0, rows ? 1 : 0, red};
    const bool same = t->exec && key.n == t->key.n && key.row == t->key.row && key.batch_stats == t->key.batch_stats &&
                      key.has_y == t->key.has_y && key.has_rows == t->key.has_rows && key.red == t->key.red;
    if (!same) {
        if (t->exec) { cudaGraphExecDestroy(t->exec); t->exec = nullptr; }
        TD_CUDA(cudaStreamBeginCapture(t->cap, cudaStreamCaptureModeThreadLocal));
        rc = td_enqueue(t, t->d_xs, y ? t->d_ys : nullptr, rows ? t->d_rows : nullptr, default_row, n, batch_stats, red, t->cap);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(t->cap, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return nf::set_error(NF_ERR_CUDA, "cudaStreamEndCapture", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&t->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { t->exec = nullptr; return nf::set_error(NF_ERR_CUDA, "cudaGraphInstantiate", cudaGetErrorString(e)); }
        t->key = key;
    }
    TD_CUDA(cudaEventRecord(t->ev_in, user));
    TD_CUDA(cudaStreamWaitEvent(t->cap, t->ev_in, 0));
    TD_CUDA(cudaGraphLaunch(t->exec, t->cap));
    TD_CUDA(cudaEventRecord(t->ev_out, t->cap));
    TD_CUDA(cudaStreamWaitEvent(user, t->ev_out, 0));
    return NF_OK;
}

int nf_trainer_apply(nf_trainer* t, const double* red, double lr, double beta1, double beta2, double eps, int world_size,
                     int update_bn, void* stream_) {
    if (!t || !red) return nf::set_error(NF_ERR_INVALID, "nf_trainer_apply", "null argument");
    if (world_size < 1) return nf::set_error(NF_ERR_INVALID, "nf_trainer_apply", "world_size must be >= 1");
    nf::td_apply_kernel<<<1, 1024, 0, (cudaStream_t)stream_>>>(t->d_prog, t->d_vars, t->d_trainable, t->d_am, t->d_av, t->d_step, red,
                                                               t->n_vars, lr, beta1, beta2, eps, 1.0 / (double)world_size, update_bn);
    TD_CUDA(cudaGetLastError());
    return NF_OK;
}

int nf_trainer_get_vars(nf_trainer* t, float* vars_host, void* stream_) {
    if (!t || !vars_host) return nf::set_error(NF_ERR_INVALID, "nf_trainer_get_vars", "null argument");
    TD_CUDA(cudaMemcpyAsync(vars_host, t->d_vars, (size_t)t->n_vars * sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
    TD_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
    return NF_OK;
}

int nf_trainer_set_vars(nf_trainer* t, const float* vars_host, void* stream_) {
    if (!t || !vars_host) return nf::set_error(NF_ERR_INVALID, "nf_trainer_set_vars", "null argument");
    TD_CUDA(cudaMemcpyAsync(t->d_vars, vars_host, (size_t)t->n_vars * sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream_));
    TD_CUDA(cudaStreamSynchronize((cudaStream_t)stream_));
    return NF_OK;
}

}  // extern "C"
