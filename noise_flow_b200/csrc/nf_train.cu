// Backward pass of the Noise Flow chain (gradient of the batch-mean NLL with respect to every trainable
// variable) for the reference's train step (train_noise_flow.py:187-198: Adam on `loss`, is_training=True).
//
// The chain is reversible layer by layer, but BatchNorm with batch statistics (layers.py:388-398) couples all
// patches of the batch, so the backward of one coupling is three passes separated by two batch-wide reductions
// (the two BatchNorm backward sums).  Activations are recomputed inside each pass from the layer's stored input
// z_in (the forward keeps every layer's input in a caller-provided workspace); nothing else is saved.
//
//   coupling forward (inverse direction), per pixel, raw (un-folded) parameters:
//     z' = z_in.A ; x0 = z'[:2], x1 = z'[2:]
//     c1 = conv3x3_SAME(x0; W1) + b1 ; h1 = relu((c1-m1)/s1) ; c2 = h1.W2 + b2 ; h2 = relu((c2-m2)/s2)
//     h3 = (conv3x3_VALID(pad(h2) (+) ring; W3) + b3) * exp(3 logs) ; shift = h3[:2], raw = h3[2:]
//     ls = scale * tanh(raw) ; out = [x0, x1*exp(ls) + shift] ; ldj = sum ls
//   loss = (1/N) sum_n [ -ldj_total_n + 0.5 sum (log 2pi + z_final^2) ]
//
//   pass B1: G_out -> g_shift, g_ls, g_x1 ; grads of scale, logs, b3, W3 ; g_h2 (transposed conv) ; BN-2 sums
//   pass B2: BN-2 backward ; grads of W2, b2 ; g_h1 ; BN-1 sums
//   pass B3: BN-1 backward ; grads of W1, b1 ; g_x0 (transposed conv) ; grad of A ; G_in = g_z' . A^T
//
// One warp owns one patch (lane = image column); images live in that warp's shared memory.  This path is written
// for clarity and exactness (fp32 math, fp64 accumulation of the parameter gradients), not for peak throughput.
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_kernels.h"
#include "nf_params.h"
#include "nf_train.h"
#include "nf_train_common.cuh"

namespace nf {

// ---------------------------------------------------------------------------------------------- pass B1
__global__ void __launch_bounds__(128, 1)
nf_train_b1_kernel(const __grid_constant__ NfTrainCoupling P, const float4* __restrict__ zin, const float4* __restrict__ gout,
                   float4* __restrict__ gzp, float4* __restrict__ scratch, long long n, float inv_n, double* __restrict__ grads) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TrainSmem& S = reinterpret_cast<TrainSmem*>(smem_raw)[warp];
    for (int k = lane; k < 34 * 34; k += 32) { S.h2[k] = make_float4(0.f, 0.f, 0.f, 0.f); S.g[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
    __syncwarp();
    float e3[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) e3[o] = expf(3.f * P.logs[o]);
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long p = (long long)blockIdx.x * (blockDim.x >> 5) + warp; p < n; p += nw) {
        load_mixed(P, S, zin + p * NF_PIXELS, lane);
        // forward net -> h2 image
        for (int r = 0; r < 32; ++r) {
            float c1hat[4], h1[4], c2hat[4];
            net_to_c2hat(P, S, r, lane, c1hat, h1, c2hat);
            S.h2[(r + 1) * 34 + lane + 1] = make_float4(fmaxf(c2hat[0], 0.f), fmaxf(c2hat[1], 0.f), fmaxf(c2hat[2], 0.f), fmaxf(c2hat[3], 0.f));
        }
        __syncwarp();
        // conv-3 forward + coupling backward at every pixel -> g_pre3 image, partial G_z'
        float g_scale = 0.f, g_logs[4] = {0.f, 0.f, 0.f, 0.f}, g_b3[4] = {0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < 32; ++r) {
            float pre[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) pre[o] = P.b3[o];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int R = r + dy, C = lane + dx;                       // padded coordinates of the tap
                    const float4 h = S.h2[R * 34 + C];
                    const float ring = (R == 0 || R == 33 || C == 0 || C == 33) ? 1.f : 0.f;
#pragma unroll
                    for (int o = 0; o < 4; ++o)
                        pre[o] += h.x * P.w3[dy][dx][0][o] + h.y * P.w3[dy][dx][1][o] + h.z * P.w3[dy][dx][2][o] +
                                  h.w * P.w3[dy][dx][3][o] + ring * P.w3[dy][dx][4][o];
                }
            float h3[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) h3[o] = pre[o] * e3[o];
            const float t0 = tanhf(h3[2]), t1 = tanhf(h3[3]);
            const float ls0 = P.scale * t0, ls1 = P.scale * t1, el0 = expf(ls0), el1 = expf(ls1);
            const float4 zp = S.zp[r * 32 + lane];
            const float4 go = gout[p * NF_PIXELS + r * 32 + lane];
            // out = [x0, x1*exp(ls)+shift]; loss has -ldj/N
            const float gls0 = go.z * zp.z * el0 - inv_n, gls1 = go.w * zp.w * el1 - inv_n;
            g_scale += gls0 * t0 + gls1 * t1;
            float gh3[4] = {go.z, go.w, gls0 * P.scale * (1.f - t0 * t0), gls1 * P.scale * (1.f - t1 * t1)};
            float gp[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) { g_logs[o] += 3.f * h3[o] * gh3[o]; gp[o] = gh3[o] * e3[o]; g_b3[o] += gp[o]; }
            S.g[(r + 1) * 34 + lane + 1] = make_float4(gp[0], gp[1], gp[2], gp[3]);
            gzp[p * NF_PIXELS + r * 32 + lane] = make_float4(go.x, go.y, go.z * el0, go.w * el1);   // x0 part completed in B3
        }
        __syncwarp();
        acc_out(grads + NF_G_SCALE, g_scale, lane);
#pragma unroll
        for (int o = 0; o < 4; ++o) { acc_out(grads + NF_G_LOGS + o, g_logs[o], lane); acc_out(grads + NF_G_B3 + o, g_b3[o], lane); }
        // grad W3[dy][dx][ci][o] = sum_pixels in(r+dy, c+dx)[ci] * g_pre3(r, c)[o]   (ci = 4: ring indicator)
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                float a[5][4];
#pragma unroll
                for (int ci = 0; ci < 5; ++ci)
#pragma unroll
                    for (int o = 0; o < 4; ++o) a[ci][o] = 0.f;
                for (int r = 0; r < 32; ++r) {
                    const int R = r + dy, C = lane + dx;
                    const float4 h = S.h2[R * 34 + C];
                    const float ring = (R == 0 || R == 33 || C == 0 || C == 33) ? 1.f : 0.f;
                    const float4 gp = S.g[(r + 1) * 34 + lane + 1];
                    const float hv[5] = {h.x, h.y, h.z, h.w, ring}, gv[4] = {gp.x, gp.y, gp.z, gp.w};
#pragma unroll
                    for (int ci = 0; ci < 5; ++ci)
#pragma unroll
                        for (int o = 0; o < 4; ++o) a[ci][o] = fmaf(hv[ci], gv[o], a[ci][o]);
                }
#pragma unroll
                for (int ci = 0; ci < 5; ++ci)
#pragma unroll
                    for (int o = 0; o < 4; ++o) acc_out(grads + NF_G_W3 + ((dy * 3 + dx) * 5 + ci) * 4 + o, a[ci][o], lane);
            }
        // g_h2(r, c)[ci] = sum_{dy,dx,o} W3[dy][dx][ci][o] * g_pre3(r-dy+1, c-dx+1)[o] ; ReLU mask ; BN-2 sums
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < 32; ++r) {
            float gh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4 gp = S.g[(r - dy + 2) * 34 + (lane - dx + 2)];   // padded index of pixel (r-dy+1, c-dx+1)
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci)
                        gh[ci] += gp.x * P.w3[dy][dx][ci][0] + gp.y * P.w3[dy][dx][ci][1] + gp.z * P.w3[dy][dx][ci][2] + gp.w * P.w3[dy][dx][ci][3];
                }
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
#pragma unroll
        for (int o = 0; o < 4; ++o) { acc_out(grads + NF_G_BN2 + o, s1[o], lane); acc_out(grads + NF_G_BN2 + 4 + o, s2[o], lane); }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------- pass B2
// bn2[0..3] = S1 / M, bn2[4..7] = S2 / M (zeros in moving-statistics mode)
__global__ void __launch_bounds__(128, 1)
nf_train_b2_kernel(const __grid_constant__ NfTrainCoupling P, const float4* __restrict__ zin, float4* __restrict__ scratch,
                   long long n, NfBnTerms bn2, double* __restrict__ grads) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TrainSmem& S = reinterpret_cast<TrainSmem*>(smem_raw)[warp];
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long p = (long long)blockIdx.x * (blockDim.x >> 5) + warp; p < n; p += nw) {
        load_mixed(P, S, zin + p * NF_PIXELS, lane);
        float gw2[4][4], gb2[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int o = 0; o < 4; ++o) gw2[i][o] = 0.f;
        for (int r = 0; r < 32; ++r) {
            float c1hat[4], h1[4], c2hat[4];
            net_to_c2hat(P, S, r, lane, c1hat, h1, c2hat);
            const float4 gc4 = scratch[p * NF_PIXELS + r * 32 + lane];
            float gc2[4], gh1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                gc2[o] = (comp(gc4, o) - bn2.v[o] - c2hat[o] * bn2.v[4 + o]) * P.is2[o];
                gb2[o] += gc2[o];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int o = 0; o < 4; ++o) { gw2[i][o] = fmaf(h1[i], gc2[o], gw2[i][o]); gh1[i] = fmaf(gc2[o], P.w2[i][o], gh1[i]); }
            float gc1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                gc1[i] = h1[i] > 0.f ? gh1[i] : 0.f;
                t1[i] += gc1[i];
                t2[i] += gc1[i] * h1[i];
            }
            scratch[p * NF_PIXELS + r * 32 + lane] = make_float4(gc1[0], gc1[1], gc1[2], gc1[3]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc_out(grads + NF_G_B2 + i, gb2[i], lane);
            acc_out(grads + NF_G_BN1 + i, t1[i], lane);
            acc_out(grads + NF_G_BN1 + 4 + i, t2[i], lane);
#pragma unroll
            for (int o = 0; o < 4; ++o) acc_out(grads + NF_G_W2 + i * 4 + o, gw2[i][o], lane);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------- pass B3
__global__ void __launch_bounds__(128, 1)
nf_train_b3_kernel(const __grid_constant__ NfTrainCoupling P, const float4* __restrict__ zin, const float4* __restrict__ scratch,
                   const float4* __restrict__ gzp, float4* __restrict__ gin, long long n, NfBnTerms bn1,
                   double* __restrict__ grads) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TrainSmem& S = reinterpret_cast<TrainSmem*>(smem_raw)[warp];
    for (int k = lane; k < 34 * 34; k += 32) S.g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long p = (long long)blockIdx.x * (blockDim.x >> 5) + warp; p < n; p += nw) {
        load_mixed(P, S, zin + p * NF_PIXELS, lane);
        float gb1[4] = {0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < 32; ++r) {
            float c1[4];
            conv1_at(P, S, r, lane, c1);
            const float4 g4 = scratch[p * NF_PIXELS + r * 32 + lane];
            float gc1[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const float c1hat = (c1[o] - P.m1[o]) * P.is1[o];
                gc1[o] = (comp(g4, o) - bn1.v[o] - c1hat * bn1.v[4 + o]) * P.is1[o];
                gb1[o] += gc1[o];
            }
            S.g[(r + 1) * 34 + lane + 1] = make_float4(gc1[0], gc1[1], gc1[2], gc1[3]);
        }
        __syncwarp();
#pragma unroll
        for (int o = 0; o < 4; ++o) acc_out(grads + NF_G_B1 + o, gb1[o], lane);
        // grad W1[dy][dx][ci][o] = sum_pixels x0(r+dy-1, c+dx-1)[ci] * g_c1(r, c)[o]
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
                float a[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                for (int r = 0; r < 32; ++r) {
                    const int rr = r + dy - 1, cc = lane + dx - 1;
                    if (rr < 0 || rr > 31 || cc < 0 || cc > 31) continue;
                    const float4 z = S.zp[rr * 32 + cc];
                    const float4 g = S.g[(r + 1) * 34 + lane + 1];
#pragma unroll
                    for (int o = 0; o < 4; ++o) { a[0][o] = fmaf(z.x, comp(g, o), a[0][o]); a[1][o] = fmaf(z.y, comp(g, o), a[1][o]); }
                }
#pragma unroll
                for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                    for (int o = 0; o < 4; ++o) acc_out(grads + NF_G_W1 + ((dy * 3 + dx) * 2 + ci) * 4 + o, a[ci][o], lane);
            }
        // g_x0 (transposed conv), complete g_z', grad A, G_in
        float gA[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int o = 0; o < 4; ++o) gA[i][o] = 0.f;
        for (int r = 0; r < 32; ++r) {
            float gx0[2] = {0.f, 0.f};
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4 g = S.g[(r - dy + 2) * 34 + (lane - dx + 2)];
#pragma unroll
                    for (int ci = 0; ci < 2; ++ci)
                        gx0[ci] += g.x * P.w1[dy][dx][ci][0] + g.y * P.w1[dy][dx][ci][1] + g.z * P.w1[dy][dx][ci][2] + g.w * P.w1[dy][dx][ci][3];
                }
            float4 gz = gzp[p * NF_PIXELS + r * 32 + lane];
            gz.x += gx0[0];
            gz.y += gx0[1];
            float4 out = gz;
            if (P.has_mix) {
                const float4 zi = zin[p * NF_PIXELS + r * 32 + lane];
                const float zv[4] = {zi.x, zi.y, zi.z, zi.w}, gv[4] = {gz.x, gz.y, gz.z, gz.w};
                float gi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int o = 0; o < 4; ++o) { gA[i][o] = fmaf(zv[i], gv[o], gA[i][o]); gi[i] = fmaf(gv[o], P.A[i][o], gi[i]); }
                out = make_float4(gi[0], gi[1], gi[2], gi[3]);
            }
            gin[p * NF_PIXELS + r * 32 + lane] = out;
        }
        if (P.has_mix) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int o = 0; o < 4; ++o) acc_out(grads + NF_G_A + i * 4 + o, gA[i][o], lane);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------- scale layers, prior
// z_out = z_in * r (sdn: r = (a*y+b)^-1/2, gain: r = 1/g).  G_in = G_out * r.
//   sdn : v = a*y+b ; d loss / d v = -0.5 * G_out * z_out / v + 0.5 / (N v) ; grad a = sum dv*y, grad b = sum dv
//   gain: grad g = sum( -G_out * z_out / g ) + 4096 / (N g)  (or 1 / (N g) for the no-sum quirk variants)
__global__ void __launch_bounds__(256)
nf_train_scale_kernel(const float4* __restrict__ zout, const float4* __restrict__ y, const float4* __restrict__ gout,
                      float4* __restrict__ gin, const int* __restrict__ rows, int default_row, long long n, float inv_n,
                      NfTrainScale T, double* __restrict__ grads) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < n; p += nw) {
        int row = rows ? rows[p] : default_row;
        row = min(max(row, 0), NF_MAX_ROWS - 1);
        const float a = T.t[row][0], b = T.t[row][1];
        float ga = 0.f, gb = 0.f;
        for (int r = 0; r < 32; ++r) {
            const long long idx = p * NF_PIXELS + r * 32 + lane;
            const float4 zo = zout[idx], go = gout[idx];
            float4 gi;
            if (T.is_sdn) {
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
                gi = make_float4(o4[0], o4[1], o4[2], o4[3]);
            } else {
                const float ginv = 1.f / a;
                ga += -(go.x * zo.x + go.y * zo.y + go.z * zo.z + go.w * zo.w) * ginv;
                gi = make_float4(go.x * ginv, go.y * ginv, go.z * ginv, go.w * ginv);
            }
            gin[idx] = gi;
        }
        ga = tw_sum(ga);
        gb = tw_sum(gb);
        if (lane == 0) {
            if (!T.is_sdn) ga += (T.full_sum ? (float)NF_DIMS : 1.f) * inv_n / a;
            atomicAdd(grads + row * 2, (double)ga);
            if (T.is_sdn) atomicAdd(grads + row * 2 + 1, (double)gb);
        }
    }
}

// stand-alone 1x1 mix: z_out = z_in . A ; grad A[i][o] = sum z_in[i] G_out[o] ; G_in = G_out . A^T
__global__ void __launch_bounds__(256)
nf_train_mix_kernel(const float4* __restrict__ zin, const float4* __restrict__ gout, float4* __restrict__ gin, long long n,
                    NfTrainMix M, double* __restrict__ grads) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < n; p += nw) {
        float gA[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int o = 0; o < 4; ++o) gA[i][o] = 0.f;
        for (int r = 0; r < 32; ++r) {
            const long long idx = p * NF_PIXELS + r * 32 + lane;
            const float4 zi = zin[idx], go = gout[idx];
            const float zv[4] = {zi.x, zi.y, zi.z, zi.w}, gv[4] = {go.x, go.y, go.z, go.w};
            float gi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int o = 0; o < 4; ++o) { gA[i][o] = fmaf(zv[i], gv[o], gA[i][o]); gi[i] = fmaf(gv[o], M.A[i][o], gi[i]); }
            gin[idx] = make_float4(gi[0], gi[1], gi[2], gi[3]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int o = 0; o < 4; ++o) acc_out(grads + i * 4 + o, gA[i][o], lane);
    }
}

// G_final = z / N  (loss = mean over patches of 0.5 * sum z^2 + const - ldj)
__global__ void nf_train_prior_kernel(const float4* __restrict__ z, float4* __restrict__ g, long long total4, float inv_n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = z[i];
        g[i] = make_float4(v.x * inv_n, v.y * inv_n, v.z * inv_n, v.w * inv_n);
    }
}

// ---------------------------------------------------------------------------------------------- launchers
static cudaError_t train_smem_attr() {
    static bool done_dev[NF_MAX_DEVICES] = {};   // per device
    bool& done = done_dev[device_slot()];
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(nf_train_b1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(TrainSmem)));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(nf_train_b2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(TrainSmem)));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(nf_train_b3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(TrainSmem)));
    if (e != cudaSuccess) return e;
    done = true;
    return cudaSuccess;
}
static unsigned train_grid(long long n, int num_sms) { long long c = (n + 3) / 4; return (unsigned)(c > num_sms ? num_sms : c); }

cudaError_t launch_train_b1(const NfTrainCoupling& P, const float* zin, const float* gout, float* gzp, float* scratch, long long n,
                            double* grads, int num_sms, cudaStream_t s) {
    cudaError_t e = train_smem_attr();
    if (e != cudaSuccess) return e;
    nf_train_b1_kernel<<<train_grid(n, num_sms), 128, 4 * sizeof(TrainSmem), s>>>(
        P, (const float4*)zin, (const float4*)gout, (float4*)gzp, (float4*)scratch, n, 1.f / (float)n, grads);
    return cudaGetLastError();
}
cudaError_t launch_train_b2(const NfTrainCoupling& P, const float* zin, float* scratch, long long n, const NfBnTerms& bn2,
                            double* grads, int num_sms, cudaStream_t s) {
    cudaError_t e = train_smem_attr();
    if (e != cudaSuccess) return e;
    nf_train_b2_kernel<<<train_grid(n, num_sms), 128, 4 * sizeof(TrainSmem), s>>>(P, (const float4*)zin, (float4*)scratch, n, bn2, grads);
    return cudaGetLastError();
}
cudaError_t launch_train_b3(const NfTrainCoupling& P, const float* zin, const float* scratch, const float* gzp, float* gin,
                            long long n, const NfBnTerms& bn1, double* grads, int num_sms, cudaStream_t s) {
    cudaError_t e = train_smem_attr();
    if (e != cudaSuccess) return e;
    nf_train_b3_kernel<<<train_grid(n, num_sms), 128, 4 * sizeof(TrainSmem), s>>>(
        P, (const float4*)zin, (const float4*)scratch, (const float4*)gzp, (float4*)gin, n, bn1, grads);
    return cudaGetLastError();
}
cudaError_t launch_train_scale(const float* zout, const float* y, const float* gout, float* gin, const int* rows, int default_row,
                               long long n, const NfTrainScale& T, double* grads, int num_sms, cudaStream_t s) {
    long long c = (n + 7) / 8;
    if (c > (long long)num_sms * 8) c = (long long)num_sms * 8;
    nf_train_scale_kernel<<<(unsigned)c, 256, 0, s>>>((const float4*)zout, (const float4*)y, (const float4*)gout, (float4*)gin, rows,
                                                     default_row, n, 1.f / (float)n, T, grads);
    return cudaGetLastError();
}
cudaError_t launch_train_mix(const float* zin, const float* gout, float* gin, long long n, const NfTrainMix& M, double* grads,
                             int num_sms, cudaStream_t s) {
    long long c = (n + 7) / 8;
    if (c > (long long)num_sms * 8) c = (long long)num_sms * 8;
    nf_train_mix_kernel<<<(unsigned)c, 256, 0, s>>>((const float4*)zin, (const float4*)gout, (float4*)gin, n, M, grads);
    return cudaGetLastError();
}
cudaError_t launch_train_prior(const float* z, float* g, long long n, int num_sms, cudaStream_t s) {
    const long long total4 = n * NF_PIXELS;
    long long c = (total4 + 255) / 256;
    if (c > (long long)num_sms * 16) c = (long long)num_sms * 16;
    nf_train_prior_kernel<<<(unsigned)c, 256, 0, s>>>((const float4*)z, (float4*)g, total4, 1.f / (float)n);
    return cudaGetLastError();
}

}  // namespace nf
