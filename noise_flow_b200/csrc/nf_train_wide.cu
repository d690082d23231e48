// Backward pass of one coupling for coupling nets wider than 4 (widths 8 / 16 / 32; `--width`, sidd/ArgParser.py:43):
// the train step of the reference (train_noise_flow.py:187-198: Adam on `loss`, is_training=True) trains any width.
// Same three passes as nf_train.cu -- the BatchNorm backward sums couple all patches of the batch --
//   B1: G_out -> g_shift, g_ls, g_x1 ; grads of scale, logs, b3, W3 ; g_h2 (transposed conv) ; BatchNorm-2 sums
//   B2: BatchNorm-2 backward ; grads of W2, b2 ; g_h1 ; BatchNorm-1 sums
//   B3: BatchNorm-1 backward ; grads of W1, b1 ; g_x0 (transposed conv) ; grad of A ; G_in = g_z' . A^T
// with activations recomputed from the layer's stored input, but ONE CTA OWNS ONE PATCH (512 threads; a thread owns the
// pixels (warp, lane) and (warp + 16, lane), as nf_wide.cu): the hidden image (34 x 34 x W) lives in shared memory as
// channel-group planes.  Parameter gradients are sums over pixels of outer products; every thread forms the products of its
// own pixels and the CTA adds them with a transposed warp butterfly (K values cost K - 1 shuffles, lane l ends up with
// value l) + shared-memory accumulators, flushed once per patch with fp64 atomics.
// Written for exactness (fp32 math, fp64 accumulation across patches), not for peak throughput.
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_kernels.h"
#include "nf_params.h"
#include "nf_train.h"

namespace nf {
namespace tw {

constexpr int THREADS = 512, WARPS = 16;

template <int W>
struct Lay {      // device parameter block of one coupling (floats), see NfTrainWideLayout in nf_train.h
    static constexpr int A = 0, META = 16, B3 = 20, LOGS = 24, B1 = 32, M1 = B1 + W, IS1 = M1 + W, B2 = IS1 + W, M2 = B2 + W, IS2 = M2 + W,
                         W1 = IS2 + W, W2 = W1 + 18 * W, W3 = W2 + W * W, SIZE = W3 + 36 * (W + 1);
};
template <int W>
struct GL {       // gradient block (doubles)
    static constexpr int A = 0, W1 = 16, B1 = W1 + 18 * W, W2 = B1 + W, B2 = W2 + W * W, W3 = B2 + W, B3 = W3 + 36 * (W + 1), LOGS = B3 + 4,
                         SCALE = LOGS + 4, BN2 = SCALE + 4, BN1 = BN2 + 2 * W, SIZE = BN1 + 2 * W;
};
static_assert(GL<8>::SIZE == nf_train_wide_grad_doubles(8) && GL<32>::SIZE == nf_train_wide_grad_doubles(32), "gradient layout");
static_assert(Lay<8>::SIZE == nf_train_wide_param_floats(8) && Lay<32>::SIZE == nf_train_wide_param_floats(32), "parameter layout");

template <int W>
struct Smem {
    static constexpr int G = W / 4;
    // accumulators -- B1: [W3 36(W+1)][scale, logs, b3: 16][BN2 2W]; B2: [W2 W*W][b2 W][BN1 2W]; B3: [W1 18W][b1 W][A 16]
    static constexpr int NACC = 36 * (W + 1) + 16 + 2 * W;
    static_assert(NACC >= W * W + 3 * W && NACC >= 19 * W + 16, "accumulator block");
    float4 zp[NF_PIXELS];              // z' = z_in . A
    float4 img[34 * G * 34];           // padded hidden image as channel-group planes: h2 (B1), g_c1 (B3); ring = 0
    float4 gimg[34 * 34];              // padded g_pre3 image (B1); ring = 0
    float w[Lay<W>::SIZE];
    float acc[NACC];
    float bn[2 * W];
};

// transposed butterfly: K values per lane -> lane l (l < K) adds the warp-wide sum of value l to acc[l]
template <int K>
__device__ __forceinline__ void reduce_add(float (&v)[K], float* acc, int lane) {
    static_assert(K == 4 || K == 8 || K == 16 || K == 32, "power of two <= 32");
#pragma unroll
    for (int off = K / 2; off > 0; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < off; ++k) {
            const float send = up ? v[k] : v[k + off], mine = up ? v[k + off] : v[k];
            v[k] = mine + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    float t = v[0];
#pragma unroll
    for (int m = K; m < 32; m <<= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
    if (lane < K) atomicAdd(acc + lane, t);
}
template <int K>
__device__ __forceinline__ void reduce_add_long(float (&v)[K], float* acc, int lane) {   // K a multiple of 32, or <= 32
    if constexpr (K <= 32) {
        reduce_add<K>(v, acc, lane);
    } else {
#pragma unroll
        for (int c = 0; c < K / 32; ++c) {
            float t[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) t[k] = v[c * 32 + k];
            reduce_add<32>(t, acc + c * 32, lane);
        }
    }
}

__device__ __forceinline__ float4 mixf(float4 v, const float* A) {   // out[o] = sum_i v[i] * A[i][o]
    float r[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) r[o] = v.x * A[0 * 4 + o] + v.y * A[1 * 4 + o] + v.z * A[2 * 4 + o] + v.w * A[3 * 4 + o];
    return make_float4(r[0], r[1], r[2], r[3]);
}

template <int W>
__device__ __forceinline__ void stage(Smem<W>& S, const float* __restrict__ params, const float4* __restrict__ zin) {
    for (int k = threadIdx.x; k < Lay<W>::SIZE; k += THREADS) S.w[k] = params[k];
    for (int k = threadIdx.x; k < Smem<W>::NACC; k += THREADS) S.acc[k] = 0.f;
    __syncthreads();
    const bool has_mix = S.w[Lay<W>::META] != 0.f;
    for (int k = threadIdx.x; k < NF_PIXELS; k += THREADS) {
        float4 z = zin[k];
        if (has_mix) z = mixf(z, S.w + Lay<W>::A);
        S.zp[k] = z;
    }
    __syncthreads();
}

// c1 (pre-BatchNorm conv-1 output) of pixel (r, c)
template <int W>
__device__ __forceinline__ void conv1_at(const Smem<W>& S, int r, int c, float (&c1)[W]) {
    using L = Lay<W>;
#pragma unroll
    for (int o = 0; o < W; ++o) c1[o] = S.w[L::B1 + o];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int rr = r + dy - 1, cc = c + dx - 1;
            if (rr < 0 || rr > 31 || cc < 0 || cc > 31) continue;
            const float4 z = S.zp[rr * 32 + cc];
            const float* w0 = S.w + L::W1 + ((dy * 3 + dx) * 2 + 0) * W;
            const float* w1 = w0 + W;
#pragma unroll
            for (int o = 0; o < W; o += 4) {
                const float4 a = *reinterpret_cast<const float4*>(w0 + o), b = *reinterpret_cast<const float4*>(w1 + o);
                c1[o] = fmaf(z.x, a.x, fmaf(z.y, b.x, c1[o]));
                c1[o + 1] = fmaf(z.x, a.y, fmaf(z.y, b.y, c1[o + 1]));
                c1[o + 2] = fmaf(z.x, a.z, fmaf(z.y, b.z, c1[o + 2]));
                c1[o + 3] = fmaf(z.x, a.w, fmaf(z.y, b.w, c1[o + 3]));
            }
        }
}
// h1 = relu(BN1(c1)) in place; c2hat = BN2(h1 . W2 + b2)
template <int W>
__device__ __forceinline__ void to_c2hat(const Smem<W>& S, float (&h1)[W], float (&c2hat)[W]) {
    using L = Lay<W>;
#pragma unroll
    for (int o = 0; o < W; ++o) h1[o] = fmaxf((h1[o] - S.w[L::M1 + o]) * S.w[L::IS1 + o], 0.f);
#pragma unroll
    for (int o = 0; o < W; ++o) c2hat[o] = S.w[L::B2 + o];
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const float* wr = S.w + L::W2 + i * W;
#pragma unroll
        for (int o = 0; o < W; o += 4) {
            const float4 a = *reinterpret_cast<const float4*>(wr + o);
            c2hat[o] = fmaf(h1[i], a.x, c2hat[o]);
            c2hat[o + 1] = fmaf(h1[i], a.y, c2hat[o + 1]);
            c2hat[o + 2] = fmaf(h1[i], a.z, c2hat[o + 2]);
            c2hat[o + 3] = fmaf(h1[i], a.w, c2hat[o + 3]);
        }
    }
#pragma unroll
    for (int o = 0; o < W; ++o) c2hat[o] = (c2hat[o] - S.w[L::M2 + o]) * S.w[L::IS2 + o];
}

template <int W>
__device__ __forceinline__ void zero_rings(Smem<W>& S) {
    constexpr int G = W / 4;
    for (int k = threadIdx.x; k < 34 * G * 34; k += THREADS) {
        const int R = k / (G * 34), C = k % 34;
        if (R == 0 || R == 33 || C == 0 || C == 33) S.img[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int k = threadIdx.x; k < 34 * 34; k += THREADS) S.gimg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
}
template <int W>
__device__ __forceinline__ void flush(Smem<W>& S, double* grads, int first, int count, int dst) {   // acc[first..] -> grads[dst..]
    for (int k = threadIdx.x; k < count; k += THREADS) {
        const float v = S.acc[first + k];
        if (v != 0.f) atomicAdd(grads + dst + k, (double)v);
        S.acc[first + k] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------- pass B1
template <int W>
__global__ void __launch_bounds__(THREADS, 1)
tw_b1_kernel(const float* __restrict__ params, const float4* __restrict__ zin, const float4* __restrict__ gout, float4* __restrict__ gzp,
             float* __restrict__ scratch, long long n, float inv_n, double* __restrict__ grads) {
    using L = Lay<W>;
    constexpr int G = W / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<W>& S = *reinterpret_cast<Smem<W>*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows2[2] = {warp, warp + 16};
    zero_rings(S);
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        stage(S, params, zin + p * NF_PIXELS);
        // forward net -> h2 planes
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            float h1[W], c2hat[W];
            conv1_at(S, rows2[q], lane, h1);
            to_c2hat(S, h1, c2hat);
#pragma unroll
            for (int g = 0; g < G; ++g)
                S.img[((rows2[q] + 1) * G + g) * 34 + lane + 1] = make_float4(fmaxf(c2hat[4 * g], 0.f), fmaxf(c2hat[4 * g + 1], 0.f),
                                                                             fmaxf(c2hat[4 * g + 2], 0.f), fmaxf(c2hat[4 * g + 3], 0.f));
        }
        __syncthreads();
        // conv-3 forward + coupling backward at the own pixels -> g_pre3 image, partial g_z'
        const float scale = S.w[L::META + 1];
        float e3[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) e3[o] = expf(3.f * S.w[L::LOGS + o]);
        float small[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) small[k] = 0.f;     // [0] g_scale, [1..4] g_logs, [5..8] g_b3
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const int r = rows2[q];
            float pre[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) pre[o] = S.w[L::B3 + o];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int R = r + dy, C = lane + dx;
                    const float* wt = S.w + L::W3 + (dy * 3 + dx) * (W + 1) * 4;
#pragma unroll 4
                    for (int g = 0; g < G; ++g) {
                        const float4 h = S.img[(R * G + g) * 34 + C];
                        const float4 w0 = *reinterpret_cast<const float4*>(wt + (4 * g) * 4), w1 = *reinterpret_cast<const float4*>(wt + (4 * g + 1) * 4),
                                     w2 = *reinterpret_cast<const float4*>(wt + (4 * g + 2) * 4), w3 = *reinterpret_cast<const float4*>(wt + (4 * g + 3) * 4);
                        pre[0] += h.x * w0.x + h.y * w1.x + h.z * w2.x + h.w * w3.x;
                        pre[1] += h.x * w0.y + h.y * w1.y + h.z * w2.y + h.w * w3.y;
                        pre[2] += h.x * w0.z + h.y * w1.z + h.z * w2.z + h.w * w3.z;
                        pre[3] += h.x * w0.w + h.y * w1.w + h.z * w2.w + h.w * w3.w;
                    }
                    if (R == 0 || R == 33 || C == 0 || C == 33) {      // edge-indicator channel (layers.py:567-571)
                        const float4 wr = *reinterpret_cast<const float4*>(wt + W * 4);
                        pre[0] += wr.x; pre[1] += wr.y; pre[2] += wr.z; pre[3] += wr.w;
                    }
                }
            float h3[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) h3[o] = pre[o] * e3[o];
            const float t0 = tanhf(h3[2]), t1 = tanhf(h3[3]);
            const float ls0 = scale * t0, ls1 = scale * t1, el0 = expf(ls0), el1 = expf(ls1);
            const float4 zp = S.zp[r * 32 + lane];
            const float4 go = gout[p * NF_PIXELS + r * 32 + lane];
            const float gls0 = go.z * zp.z * el0 - inv_n, gls1 = go.w * zp.w * el1 - inv_n;      // loss has -ldj / N
            small[0] += gls0 * t0 + gls1 * t1;
            const float gh3[4] = {go.z, go.w, gls0 * scale * (1.f - t0 * t0), gls1 * scale * (1.f - t1 * t1)};
            float gp[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) { small[1 + o] += 3.f * h3[o] * gh3[o]; gp[o] = gh3[o] * e3[o]; small[5 + o] += gp[o]; }
            S.gimg[(r + 1) * 34 + lane + 1] = make_float4(gp[0], gp[1], gp[2], gp[3]);
            gzp[p * NF_PIXELS + r * 32 + lane] = make_float4(go.x, go.y, go.z * el0, go.w * el1);   // x0 part completed in B3
        }
        reduce_add<16>(small, S.acc + 36 * (W + 1), lane);
        __syncthreads();
        // grad W3[tap][i][o] = sum_pixels h2pad(r + dy, c + dx)[i] * g_pre3(r, c)[o]   (i = W: edge indicator)
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;
            float4 gq[2];
            bool ring[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                gq[q] = S.gimg[(rows2[q] + 1) * 34 + lane + 1];
                const int R = rows2[q] + dy, C = lane + dx;
                ring[q] = R == 0 || R == 33 || C == 0 || C == 33;
            }
#pragma unroll 2
            for (int g = 0; g < G; ++g) {
                float v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = 0.f;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float4 h = S.img[((rows2[q] + dy) * G + g) * 34 + lane + dx];
                    const float hv[4] = {h.x, h.y, h.z, h.w}, gv[4] = {gq[q].x, gq[q].y, gq[q].z, gq[q].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int o = 0; o < 4; ++o) v[k * 4 + o] = fmaf(hv[k], gv[o], v[k * 4 + o]);
                }
                reduce_add<16>(v, S.acc + (tap * (W + 1) + 4 * g) * 4, lane);
            }
            float vr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (ring[q]) { vr[0] += gq[q].x; vr[1] += gq[q].y; vr[2] += gq[q].z; vr[3] += gq[q].w; }
            reduce_add<4>(vr, S.acc + (tap * (W + 1) + W) * 4, lane);
        }
        // g_h2(r, c)[i] = sum_{taps, o} W3[tap][i][o] * g_pre3(r - dy + 1, c - dx + 1)[o] ; ReLU mask ; BatchNorm-2 sums
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const int r = rows2[q];
            float gh[W];
#pragma unroll
            for (int i = 0; i < W; ++i) gh[i] = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4 gp = S.gimg[(r - dy + 2) * 34 + (lane - dx + 2)];      // padded index of pixel (r - dy + 1, c - dx + 1)
                    const float* wt = S.w + L::W3 + (dy * 3 + dx) * (W + 1) * 4;
#pragma unroll
                    for (int i = 0; i < W; ++i) {
                        const float4 wv = *reinterpret_cast<const float4*>(wt + i * 4);
                        gh[i] += gp.x * wv.x + gp.y * wv.y + gp.z * wv.z + gp.w * wv.w;
                    }
                }
            float s2[W];
            float* dst = scratch + ((size_t)p * NF_PIXELS + r * 32 + lane) * W;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float4 h = S.img[((r + 1) * G + g) * 34 + lane + 1];
                const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = 4 * g + k;
                    gh[i] = hv[k] > 0.f ? gh[i] : 0.f;
                    s2[i] = gh[i] * hv[k];          // = g_c2hat * c2hat wherever the mask is on
                }
                *reinterpret_cast<float4*>(dst + 4 * g) = make_float4(gh[4 * g], gh[4 * g + 1], gh[4 * g + 2], gh[4 * g + 3]);
            }
            reduce_add_long<W>(gh, S.acc + 36 * (W + 1) + 16, lane);
            reduce_add_long<W>(s2, S.acc + 36 * (W + 1) + 16 + W, lane);
        }
        __syncthreads();
        flush(S, grads, 0, 36 * (W + 1), GL<W>::W3);
        flush(S, grads, 36 * (W + 1), 1, GL<W>::SCALE);
        flush(S, grads, 36 * (W + 1) + 1, 4, GL<W>::LOGS);
        flush(S, grads, 36 * (W + 1) + 5, 4, GL<W>::B3);
        flush(S, grads, 36 * (W + 1) + 16, 2 * W, GL<W>::BN2);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- pass B2
// bn[0..W) = S1 / M, bn[W..2W) = S2 / M of BatchNorm-2 (zeros in moving-statistics mode)
template <int W>
__global__ void __launch_bounds__(THREADS, 1)
tw_b2_kernel(const float* __restrict__ params, const float4* __restrict__ zin, float* __restrict__ scratch, long long n,
             const NfBnTermsWide bn, double* __restrict__ grads) {
    using L = Lay<W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<W>& S = *reinterpret_cast<Smem<W>*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows2[2] = {warp, warp + 16};
    for (int k = threadIdx.x; k < 2 * W; k += THREADS) S.bn[k] = bn.v[k];
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        stage(S, params, zin + p * NF_PIXELS);
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const int r = rows2[q];
            float h1[W], c2hat[W];
            conv1_at(S, r, lane, h1);
            to_c2hat(S, h1, c2hat);
            float* gs = scratch + ((size_t)p * NF_PIXELS + r * 32 + lane) * W;
            float gc2[W];
#pragma unroll
            for (int o = 0; o < W; o += 4) {
                const float4 g4 = *reinterpret_cast<const float4*>(gs + o);
                const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) gc2[o + k] = (gv[k] - S.bn[o + k] - c2hat[o + k] * S.bn[W + o + k]) * S.w[L::IS2 + o + k];
            }
            // grad W2[i][o] = sum_pixels h1[i] * g_c2[o]
#pragma unroll 1
            for (int i = 0; i < W; ++i) {
                float hi = 0.f;
#pragma unroll
                for (int k = 0; k < W; ++k) hi = k == i ? h1[k] : hi;      // h1[i] with a run-time i, registers only
                float v[W];
#pragma unroll
                for (int o = 0; o < W; ++o) v[o] = hi * gc2[o];
                reduce_add_long<W>(v, S.acc + i * W, lane);
            }
            // g_h1[i] = sum_o W2[i][o] * g_c2[o] ; ReLU mask ; BatchNorm-1 sums
            float gc1[W], t2[W];
#pragma unroll
            for (int i = 0; i < W; ++i) {
                const float* wr = S.w + L::W2 + i * W;
                float a = 0.f;
#pragma unroll
                for (int o = 0; o < W; o += 4) {
                    const float4 wv = *reinterpret_cast<const float4*>(wr + o);
                    a += gc2[o] * wv.x + gc2[o + 1] * wv.y + gc2[o + 2] * wv.z + gc2[o + 3] * wv.w;
                }
                gc1[i] = h1[i] > 0.f ? a : 0.f;
                t2[i] = gc1[i] * h1[i];
            }
#pragma unroll
            for (int o = 0; o < W; o += 4) *reinterpret_cast<float4*>(gs + o) = make_float4(gc1[o], gc1[o + 1], gc1[o + 2], gc1[o + 3]);
            reduce_add_long<W>(gc2, S.acc + W * W, lane);
            reduce_add_long<W>(gc1, S.acc + W * W + W, lane);
            reduce_add_long<W>(t2, S.acc + W * W + 2 * W, lane);
        }
        __syncthreads();
        flush(S, grads, 0, W * W, GL<W>::W2);
        flush(S, grads, W * W, W, GL<W>::B2);
        flush(S, grads, W * W + W, 2 * W, GL<W>::BN1);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- pass B3
template <int W>
__global__ void __launch_bounds__(THREADS, 1)
tw_b3_kernel(const float* __restrict__ params, const float4* __restrict__ zin, const float* __restrict__ scratch,
             const float4* __restrict__ gzp, float4* __restrict__ gin, long long n, const NfBnTermsWide bn, double* __restrict__ grads) {
    using L = Lay<W>;
    constexpr int G = W / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<W>& S = *reinterpret_cast<Smem<W>*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows2[2] = {warp, warp + 16};
    zero_rings(S);
    for (int k = threadIdx.x; k < 2 * W; k += THREADS) S.bn[k] = bn.v[k];
    for (long long p = blockIdx.x; p < n; p += gridDim.x) {
        stage(S, params, zin + p * NF_PIXELS);
        const bool has_mix = S.w[L::META] != 0.f;
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const int r = rows2[q];
            float c1[W];
            conv1_at(S, r, lane, c1);
            const float* gs = scratch + ((size_t)p * NF_PIXELS + r * 32 + lane) * W;
            float gc1[W];
#pragma unroll
            for (int o = 0; o < W; o += 4) {
                const float4 g4 = *reinterpret_cast<const float4*>(gs + o);
                const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float c1hat = (c1[o + k] - S.w[L::M1 + o + k]) * S.w[L::IS1 + o + k];
                    gc1[o + k] = (gv[k] - S.bn[o + k] - c1hat * S.bn[W + o + k]) * S.w[L::IS1 + o + k];
                }
                S.img[((r + 1) * G + (o >> 2)) * 34 + lane + 1] = make_float4(gc1[o], gc1[o + 1], gc1[o + 2], gc1[o + 3]);
            }
            // grad W1[tap][i][o] = sum_pixels x0(r + dy - 1, c + dx - 1)[i] * g_c1(r, c)[o]
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
                const int rr = r + tap / 3 - 1, cc = lane + tap % 3 - 1;
                float2 x0 = make_float2(0.f, 0.f);
                if (rr >= 0 && rr <= 31 && cc >= 0 && cc <= 31) { const float4 z = S.zp[rr * 32 + cc]; x0 = make_float2(z.x, z.y); }
                float v[W];
#pragma unroll
                for (int o = 0; o < W; ++o) v[o] = x0.x * gc1[o];
                reduce_add_long<W>(v, S.acc + (tap * 2 + 0) * W, lane);
#pragma unroll
                for (int o = 0; o < W; ++o) v[o] = x0.y * gc1[o];
                reduce_add_long<W>(v, S.acc + (tap * 2 + 1) * W, lane);
            }
            reduce_add_long<W>(gc1, S.acc + 18 * W, lane);
        }
        __syncthreads();
        // g_x0 (transposed conv), complete g_z', grad A, G_in
        float gA[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) gA[k] = 0.f;
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
            const int r = rows2[q];
            float gx0[2] = {0.f, 0.f};
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float* w0 = S.w + L::W1 + ((dy * 3 + dx) * 2 + 0) * W;
                    const float* w1 = w0 + W;
#pragma unroll 4
                    for (int g = 0; g < G; ++g) {
                        const float4 gq = S.img[((r - dy + 2) * G + g) * 34 + (lane - dx + 2)];
                        const float4 a = *reinterpret_cast<const float4*>(w0 + 4 * g), b = *reinterpret_cast<const float4*>(w1 + 4 * g);
                        gx0[0] += gq.x * a.x + gq.y * a.y + gq.z * a.z + gq.w * a.w;
                        gx0[1] += gq.x * b.x + gq.y * b.y + gq.z * b.z + gq.w * b.w;
                    }
                }
            float4 gz = gzp[p * NF_PIXELS + r * 32 + lane];
            gz.x += gx0[0];
            gz.y += gx0[1];
            float4 out = gz;
            if (has_mix) {
                const float4 zi = zin[p * NF_PIXELS + r * 32 + lane];
                const float zv[4] = {zi.x, zi.y, zi.z, zi.w}, gv[4] = {gz.x, gz.y, gz.z, gz.w};
                float gi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int o = 0; o < 4; ++o) { gA[i * 4 + o] = fmaf(zv[i], gv[o], gA[i * 4 + o]); gi[i] = fmaf(gv[o], S.w[L::A + i * 4 + o], gi[i]); }
                out = make_float4(gi[0], gi[1], gi[2], gi[3]);
            }
            gin[p * NF_PIXELS + r * 32 + lane] = out;
        }
        reduce_add<16>(gA, S.acc + 19 * W, lane);
        __syncthreads();
        flush(S, grads, 0, 18 * W, GL<W>::W1);
        flush(S, grads, 18 * W, W, GL<W>::B1);
        if (has_mix) flush(S, grads, 19 * W, 16, GL<W>::A);
        __syncthreads();
    }
}

template <int W>
static cudaError_t attrs() {
    static bool done_dev[NF_MAX_DEVICES] = {};
    bool& done = done_dev[device_slot()];
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(tw_b1_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<W>));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tw_b2_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<W>));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tw_b3_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<W>));
    done = e == cudaSuccess;
    return e;
}

template <int W>
static cudaError_t run(int pass, const float* params, const float* zin, const float* gout, float* gzp, float* scratch, float* gin, long long n,
                       const NfBnTermsWide* bn, double* grads, int num_sms, cudaStream_t s) {
    cudaError_t e = attrs<W>();
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)(n < (long long)num_sms ? n : (long long)num_sms);
    const size_t smem = sizeof(Smem<W>);
    if (pass == 1)
        tw_b1_kernel<W><<<grid, THREADS, smem, s>>>(params, (const float4*)zin, (const float4*)gout, (float4*)gzp, scratch, n, 1.f / (float)n, grads);
    else if (pass == 2)
        tw_b2_kernel<W><<<grid, THREADS, smem, s>>>(params, (const float4*)zin, scratch, n, *bn, grads);
    else
        tw_b3_kernel<W><<<grid, THREADS, smem, s>>>(params, (const float4*)zin, scratch, (const float4*)gzp, (float4*)gin, n, *bn, grads);
    return cudaGetLastError();
}

}  // namespace tw

bool train_wide_width_supported(int W) { return W == 8 || W == 16 || W == 32; }

cudaError_t launch_train_wide(int W, int pass, const float* params, const float* zin, const float* gout, float* gzp, float* scratch, float* gin,
                              long long n, const NfBnTermsWide* bn, double* grads, int num_sms, cudaStream_t s) {
    switch (W) {
        case 8: return tw::run<8>(pass, params, zin, gout, gzp, scratch, gin, n, bn, grads, num_sms, s);
        case 16: return tw::run<16>(pass, params, zin, gout, gzp, scratch, gin, n, bn, grads, num_sms, s);
        case 32: return tw::run<32>(pass, params, zin, gout, gzp, scratch, gin, n, bn, grads, num_sms, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace nf
