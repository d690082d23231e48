// One affine coupling (+ fused 1x1 mix) as a single software-pipelined pass over the 32 rows of a patch.
// Included by nf_kernels.cu (needs WarpSmem, ffma2, ld2, mix4, fast_tanh, fast_exp from there).
//
// Row schedule (t = 0..35); every stage of a step works on data published in EARLIER steps, so a step
// needs exactly one __syncwarp:
//   stage A (t     = 0..31): z[t] <- z[t].A (inverse only); publish x0 row t (channels 0,1) to ring slot t&1
//   stage B (i=t-1 = 0..31): scatter x0 row i into the pending conv-1 rows: finish row i-1 (dy = 2), continue
//                            row i (dy = 1), start row i+1 (dy = 0); the finished row goes through
//                            BN+ReLU, 1x1 conv, BN+ReLU and is published as h2 row i-1
//   stage C (j=t-3 = 0..31): same scatter of h2 row j into the pending conv-3 rows; the finished row q = j-1
//                            gets the edge-indicator bias, then the tanh/exp affine update of z[q] + log-det
// Only two pending rows per convolution are carried between steps (old = row i partial, mid = row i+1
// partial).  The scatter is written out-of-place -- fin = old + W2.in, old' = mid + W1.in, mid' = W0.in --
// so the role rotation is pure register renaming (the first FMA of each chain writes the destination
// role), no MOVs.  Steps 5..31 are the guard-free steady state; prologue / epilogue steps share one
// guarded copy of the body.
#pragma once

namespace nf {

struct B3 { float top[4], mid[4], bot[4]; };   // conv2d_zeros bias incl. edge indicator, for this lane's column

template <class CP>
__device__ __forceinline__ B3 load_b3(const CP& P, int lane) {
    B3 b;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        b.top[o] = lane == 0 ? P.b3[0][0][o] : (lane == 31 ? P.b3[0][2][o] : P.b3[0][1][o]);
        b.mid[o] = lane == 0 ? P.b3[1][0][o] : (lane == 31 ? P.b3[1][2][o] : P.b3[1][1][o]);
        b.bot[o] = lane == 0 ? P.b3[2][0][o] : (lane == 31 ? P.b3[2][2][o] : P.b3[2][1][o]);
    }
    return b;
}

// STATS: 0 = normal step.  1 / 2 = batch-statistics probe (reference batch_norm(training=True),
// layers.py:388-398): accumulate sum and sum-of-squares of the conv-1 (1) or conv-2 (2) output BEFORE its
// BatchNorm into stats[0..3] / stats[4..7]; nothing is published and stage C does not run.  The caller folds
// the BatchNorm under probe as the identity, so the folded activation IS the raw pre-BN activation.
template <bool INV, bool GUARDED, int STATS = 0, class CP>
__device__ __forceinline__ void coupling_step(const CP& P, WarpSmem& s, const ZStore& zs, const int lane, const int t,
                                              const bool has_mix, Acc4& b_old, Acc4& b_mid, Acc4& c_old, Acc4& c_mid,
                                              const B3& b3, float& ldj, const float2 (&am)[4][2], const float2 (&w2c)[4][2], float* stats = nullptr) {
    const float2 zero2 = make_float2(0.f, 0.f);
    constexpr bool W2R = !INV && STATS == 0;
    const bool do_a = !GUARDED || t < 32;
    const bool b_fma = !GUARDED || (t >= 1 && t <= 32);
    const bool b_emit = !GUARDED || (t >= 2 && t <= 33);
    const bool c_fma = STATS == 0 && (!GUARDED || (t >= 3 && t <= 34));
    const bool c_emit = STATS == 0 && (!GUARDED || t >= 4);
    // z rows of this step: row t (stage A) and row t-4 (stage C), fetched together (one wait for both).
    // commit() (= tcgen05.wait::st when z lives in TMEM) sits HERE, not after the stores: the stores of the previous
    // step were issued a few hundred cycles ago, so the wait is free, and row t-4 was last written 4 steps back.
    zs.commit();
    float4 za = make_float4(0.f, 0.f, 0.f, 0.f), zc = za;
    if (do_a && c_emit) zs.load2(t, t - 4, za, zc);
    else if (do_a) za = zs.load(t);
    else if (c_emit) zc = zs.load(t - 4);
    // ---------------- stage A
    if (do_a) {
        float4 z = za;
        if (INV && has_mix) {
            z = mix4r(z, am);                                   // Conv2d1x1._inverse, layers.py:117-119
            zs.store(t, z);
        }
        s.xr[t & 1][lane + 1] = make_float2(z.x, z.y);
    }
    // ---------------- stage B
    Acc4 fin = b_old;
    if (b_fma) {
        const int i = t - 1;
        float2 xin[3];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) xin[dx] = s.xr[i & 1][lane + dx];
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
        for (int o = 0; o < 4; ++o) {
            const float c1 = fin.v[o].x + fin.v[o].y + P.b1[o];
            if (STATS == 1) { stats[o] += c1; stats[4 + o] = fmaf(c1, c1, stats[4 + o]); }
            h1[o] = fmaxf(c1, 0.f);                                                          // BN folded, ReLU
        }
        const float2 h01 = make_float2(h1[0], h1[1]), h23 = make_float2(h1[2], h1[3]);
        float h2[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float2 u = ffma2(h01, W2R ? w2c[o][0] : ld2(&P.w2[o][0]), zero2);
            u = ffma2(h23, W2R ? w2c[o][1] : ld2(&P.w2[o][2]), u);
            const float c2 = u.x + u.y + P.b2[o];
            if (STATS == 2) { stats[o] += c2; stats[4 + o] = fmaf(c2, c2, stats[4 + o]); }
            h2[o] = fmaxf(c2, 0.f);
        }
        if (STATS == 0) s.hr[r & 1][lane + 1] = make_float4(h2[0], h2[1], h2[2], h2[3]);
    }
    // ---------------- stage C
    Acc4 cfin = c_old;
    if (c_fma) {
        const int j = t - 3;
        float4 hin[3];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) hin[dx] = s.hr[j & 1][lane + dx];
        Acc4 nold, nmid;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const float2 lo = make_float2(hin[dx].x, hin[dx].y), hi = make_float2(hin[dx].z, hin[dx].w);
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                cfin.v[o] = ffma2(lo, ld2(&P.w3[2][dx][o][0]), dx == 0 ? c_old.v[o] : cfin.v[o]);
                nold.v[o] = ffma2(lo, ld2(&P.w3[1][dx][o][0]), dx == 0 ? c_mid.v[o] : nold.v[o]);
                nmid.v[o] = ffma2(lo, ld2(&P.w3[0][dx][o][0]), dx == 0 ? zero2 : nmid.v[o]);
                cfin.v[o] = ffma2(hi, ld2(&P.w3[2][dx][o][2]), cfin.v[o]);
                nold.v[o] = ffma2(hi, ld2(&P.w3[1][dx][o][2]), nold.v[o]);
                nmid.v[o] = ffma2(hi, ld2(&P.w3[0][dx][o][2]), nmid.v[o]);
            }
        }
        c_old = nold;
        c_mid = nmid;
    } else if (GUARDED) {
        c_old = c_mid;   // t = 35
    }
    if (c_emit) {
        const int q = t - 4;
        float h3[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const float b = !GUARDED ? b3.mid[o] : (q == 0 ? b3.top[o] : (q == 31 ? b3.bot[o] : b3.mid[o]));
            h3[o] = cfin.v[o].x + cfin.v[o].y + b;
        }
        // shift = h3[0:2], log_scale = scale * tanh(h3[2:4])                    (layers.py:362 / :342)
        const float ls0 = P.scale * fast_tanh(h3[2]);
        const float ls1 = P.scale * fast_tanh(h3[3]);
        float4 z = zc;
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
__device__ __forceinline__ void coupling_pass(const CP& P, WarpSmem& s, const ZStore& zs, const int lane, float& ldj) {
    const bool has_mix = P.has_mix != 0;
    const B3 b3 = load_b3(P, lane);
    float2 am[4][2];
    load_mix_regs<INV>(P, am, s.xr[0][0].x);
    float2 w2c[4][2];
    if (!INV) {
        const float rz = s.xr[0][0].x;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            w2c[o][0] = make_float2(P.w2[o][0] + rz, P.w2[o][1] + rz);
            w2c[o][1] = make_float2(P.w2[o][2] + rz, P.w2[o][3] + rz);
        }
    }
    Acc4 b_old, b_mid, c_old, c_mid;
#pragma unroll
    for (int o = 0; o < 4; ++o) b_old.v[o] = b_mid.v[o] = c_old.v[o] = c_mid.v[o] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int t = 0; t < 36; ++t) {
        if (t >= 5 && t < 32) coupling_step<INV, false>(P, s, zs, lane, t, has_mix, b_old, b_mid, c_old, c_mid, b3, ldj, am, w2c);
        else                  coupling_step<INV, true>(P, s, zs, lane, t, has_mix, b_old, b_mid, c_old, c_mid, b3, ldj, am, w2c);
    }
}

// Batch-statistics probe of one coupling: rows 0..31 of the conv-1 / conv-2 pre-BN activation.
template <bool INV, int STAGE, class CP>
__device__ __forceinline__ void coupling_stats_pass(const CP& P, WarpSmem& s, const ZStore& zs, const int lane, float* stats) {
    const bool has_mix = P.has_mix != 0;
    B3 b3 = {};
    Acc4 b_old, b_mid, c_old, c_mid;
#pragma unroll
    for (int o = 0; o < 4; ++o) b_old.v[o] = b_mid.v[o] = c_old.v[o] = c_mid.v[o] = make_float2(0.f, 0.f);
    float ldj = 0.f;
    float2 am[4][2];
    load_mix_regs<INV>(P, am, s.xr[0][0].x);
    float2 w2c[4][2] = {};
#pragma unroll 1
    for (int t = 0; t < 34; ++t)
        coupling_step<INV, true, STAGE>(P, s, zs, lane, t, has_mix, b_old, b_mid, c_old, c_mid, b3, ldj, am, w2c, stats);
}

}  // namespace nf
