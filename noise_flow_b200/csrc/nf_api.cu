// C-ABI of the B200 Noise Flow engine (declared in include/noiseflow_b200.h): model handle, host-side
// parameter folding (reference-shaped weights -> kernel-ready NfModelParams), launches, host-buffer
// pipeline.  No torch types, no exceptions across the boundary, no global mutable state except the
// thread-local error string.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <algorithm>
#include <mutex>
#include <utility>
#include <new>
#include <vector>

#include "../../include/noiseflow_b200.h"
#include "nf_kernels.h"
#include "nf_params.h"
#include "nf_train.h"
#include "nf_wide.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define NF_CUDA(call)                                                                                 \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) return fail(NF_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));   \
    } while (0)

enum LayerKind { L_CONV1X1 = 1, L_PERMUTE = 2, L_COUPLING = 3, L_SCALE = 4 };

struct Layer {
    int kind = 0;
    // conv1x1 / permute: mixing matrices as the kernel wants them, [o][i]
    float a[4][4] = {};
    float ainv[4][4] = {};
    double log_abs_det = 0.0;   // per pixel (conv1x1), 0 for permutations
    // coupling (folded)
    NfCouplingP cp = {};
    // reference-shaped copy, kept so that BatchNorm can be re-folded with batch statistics (is_training)
    struct Raw {
        float l1_w[72], l1_b[4], bn1_mean[4], bn1_var[4], l2_w[16], l2_b[4], bn2_mean[4], bn2_var[4];
        float last_w[180], last_b[4], last_logs[4], rescaling_scale, bn_eps;
    } raw = {};
    // coupling of a wide net (width != 4): reference-shaped copy, [l1_w 18W][l1_b W][bn1_mean W][bn1_var W][l2_w W*W]
    // [l2_b W][bn2_mean W][bn2_var W][last_w 36(W+1)][last_b 4][last_logs 4]; rescaling_scale / bn_eps in `raw`
    std::vector<float> wraw;
    // coupling kind (NF_COUPLING_MODE_*): 0 AffineCoupling, 1 CondXY[G], 2 CondY[G]; input / output channels of its net
    int cmode = 0, cin = 2, cout = 4;
    // scale
    int scale_kind = 0, full_sum = 1, n_rows = 0;
    float table[NF_MAX_ROWS][4] = {};
};

struct WideRawView {   // sub-arrays of Layer::wraw
    const float *l1_w, *l1_b, *bn1_mean, *bn1_var, *l2_w, *l2_b, *bn2_mean, *bn2_var, *last_w, *last_b, *last_logs;
    WideRawView(const float* p, int W, int cin = 2, int cout = 4) {
        l1_w = p; p += 9 * cin * W; l1_b = p; p += W; bn1_mean = p; p += W; bn1_var = p; p += W;
        l2_w = p; p += W * W; l2_b = p; p += W; bn2_mean = p; p += W; bn2_var = p; p += W;
        last_w = p; p += 9 * cout * (W + 1); last_b = p; p += cout; last_logs = p;
    }
    static size_t floats(int W, int cin = 2, int cout = 4) {
        return (size_t)9 * cin * W + 3 * W + (size_t)W * W + 3 * W + 9 * (size_t)cout * (W + 1) + 2 * cout;
    }
};

}  // namespace

struct nf_model {
    std::vector<Layer> layers;
    bool finalized = false;
    bool has_cond = false;       // clean-image-conditioned couplings (legacy revnet2d models): every launch takes the generic
                                 // CTA-per-patch kernel (nf_wide_cond.cu), at width 4 too
    int warps_per_cta = NF_MAX_WARPS_PER_CTA;
    int use_tc = 0;              // width 4, which chain kernel runs calls without batch-statistics probes:
                                 //   0 (default) = 4 = the all-fp32 vertical-Winograd kernel (nf_wino.cu: +10 % data -> latent, +3 % latent -> data);
                                 //   5 = the all-fp32 direct-form kernel (nf_kernels.cu) everywhere;
                                 //   2 = hybrid kernel (nf_hybrid.cu: conv-3 on tcgen05) everywhere; 3 = hybrid for latent -> data
                                 //       (sampling +7 % over the direct form), default otherwise;  1 = both 3x3 convs as bf16 implicit GEMMs (nf_tc.cu)
    int bs_small = 1;            // 1: nf_chain_batch_stats runs small batches as one cooperative kernel (nf_model_set_bs_small)
    // parameter-image / statistics buffers of the small-batch chain: a call takes one (or allocates it) and gives it back
    // after its stream synchronisation, so concurrent callers never share one and no call pays for cudaMalloc
    mutable std::mutex bs_mu;
    mutable std::vector<std::pair<unsigned char*, size_t>> bs_free;
    int num_ctas = 0;
    int sm_count = 0;
    NfModelParams full = {};     // fused program of the whole chain
    float full_ldj_const = 0.f;
    // _host entry points: every call borrows a HostPipe (staging slots + streams) from this free list and gives it back at
    // the end, so concurrent host callers (the reference's 16-32 sampler threads, train_dncnn_noiseflow.py:195-198) each
    // run their own copy/compute pipeline; a pipe's slots are sized by the largest call it has served (<= kChunk patches)
    std::mutex pool_mu;          // guards pipes_free only
    struct Staging {
        float *x = nullptr, *y = nullptr, *z = nullptr, *nll = nullptr, *sdz = nullptr;
        int32_t* rows = nullptr;
        int64_t cap = 0;         // patches this slot's buffers hold
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
    };
    static constexpr int kSlots = 4;
    struct HostPipe {
        Staging st[kSlots];
        float* h_tmp = nullptr;      // pinned scratch for per-patch results the caller did not ask for (sums only)
        size_t h_tmp_floats = 0;
    };
    std::vector<HostPipe*> pipes_free;
    int pipes_total = 0;
    mutable std::mutex prog_mu;  // guards `layers` / `full` against a concurrent nf_model_set_* (launches snapshot)
    // wide coupling nets (width 8 / 16 / 32, nf_wide.cu): the folded program of the whole chain lives in a device
    // blob owned by the handle (re-uploaded by nf_model_finalize / nf_model_set_*)
    int width = 4;
    NfWideProgram wide_full = {};
    float* d_wide_full = nullptr;
    size_t wide_full_floats = 0;
    // ... and, for widths the tensor-core kernel covers (32 / 64 / 128, nf_wide_tc.cu), the same chain folded into bf16
    // (hi, lo) UMMA operand blocks.  use_tc_wide = 0 forces the CUDA-core kernel where one exists (width 32).
    NfWideProgram wide_tc_full = {};
    float* d_wide_tc = nullptr;
    int use_tc_wide = 1;
    int defer_finalize = 0;      // nf_model_begin_update .. nf_model_end_update: set many layers, fold / upload once
};

namespace {

// ---- folding: reference-shaped coupling weights -> NfCouplingP (double precision on the host) -----
// real_nvp_conv_template, layers.py:469-494 with batch_norm in moving-statistics mode (:400):
//   h1 = relu((conv(x0;W1)+b1 - m1)/sqrt(v1+eps))  ->  W1' = W1*s1, b1' = (b1-m1)*s1
//   h2 = relu((h1.W2+b2 - m2)/sqrt(v2+eps))        ->  W2' = W2*s2, b2' = (b2-m2)*s2
//   h3 = (conv_valid([pad(h2), ring]; W3) + b3) * exp(3*logs)
//      -> W3' = W3[:,:,:4,:]*e, b3'[row class][col class] = (b3 + sum of ring taps)*e
int fold_coupling(const nf_coupling_weights* w, NfCouplingP* out) {
    if (!w || !w->l1_w || !w->l1_b || !w->bn1_mean || !w->bn1_var || !w->l2_w || !w->l2_b || !w->bn2_mean ||
        !w->bn2_var || !w->last_w || !w->last_b || !w->last_logs)
        return fail(NF_ERR_INVALID, "nf_coupling_weights: null pointer");
    const double eps = w->bn_eps;
    double s1[4], s2[4], e[4];
    for (int o = 0; o < 4; ++o) {
        if (!(w->bn1_var[o] + eps > 0.0) || !(w->bn2_var[o] + eps > 0.0))
            return fail(NF_ERR_INVALID, "batch-norm variance + eps must be positive");
        s1[o] = 1.0 / sqrt((double)w->bn1_var[o] + eps);
        s2[o] = 1.0 / sqrt((double)w->bn2_var[o] + eps);
        e[o] = exp(3.0 * (double)w->last_logs[o]);
    }
    for (int dy = 0; dy < 3; ++dy)
        for (int dx = 0; dx < 3; ++dx)
            for (int o = 0; o < 4; ++o) {
                for (int i = 0; i < 2; ++i)
                    out->w1[dy][dx][o][i] = (float)((double)w->l1_w[((dy * 3 + dx) * 2 + i) * 4 + o] * s1[o]);
                for (int i = 0; i < 4; ++i)
                    out->w3[dy][dx][o][i] = (float)((double)w->last_w[((dy * 3 + dx) * 5 + i) * 4 + o] * e[o]);
            }
    for (int o = 0; o < 4; ++o) {
        out->b1[o] = (float)(((double)w->l1_b[o] - (double)w->bn1_mean[o]) * s1[o]);
        out->b2[o] = (float)(((double)w->l2_b[o] - (double)w->bn2_mean[o]) * s2[o]);
        for (int i = 0; i < 4; ++i) out->w2[o][i] = (float)((double)w->l2_w[i * 4 + o] * s2[o]);
    }
    // edge indicator (layers.py:567-571): 1 on the one-pixel ring of the 34x34 padded map.  Output
    // pixel (r,c) sees padded positions (r+dy, c+dx): on the ring iff r+dy in {0,33} or c+dx in {0,33}.
    for (int rc = 0; rc < 3; ++rc)
        for (int cc = 0; cc < 3; ++cc)
            for (int o = 0; o < 4; ++o) {
                double acc = w->last_b[o];
                for (int dy = 0; dy < 3; ++dy)
                    for (int dx = 0; dx < 3; ++dx) {
                        const bool ring = (rc == 0 && dy == 0) || (rc == 2 && dy == 2) || (cc == 0 && dx == 0) ||
                                          (cc == 2 && dx == 2);
                        if (ring) acc += (double)w->last_w[((dy * 3 + dx) * 5 + 4) * 4 + o];
                    }
                out->b3[rc][cc][o] = (float)(acc * e[o]);
            }
    out->scale = w->rescaling_scale;
    return NF_OK;
}

void keep_raw(Layer& L, const nf_coupling_weights* w) {
    memcpy(L.raw.l1_w, w->l1_w, sizeof(L.raw.l1_w));       memcpy(L.raw.l1_b, w->l1_b, sizeof(L.raw.l1_b));
    memcpy(L.raw.bn1_mean, w->bn1_mean, 16);               memcpy(L.raw.bn1_var, w->bn1_var, 16);
    memcpy(L.raw.l2_w, w->l2_w, sizeof(L.raw.l2_w));       memcpy(L.raw.l2_b, w->l2_b, sizeof(L.raw.l2_b));
    memcpy(L.raw.bn2_mean, w->bn2_mean, 16);               memcpy(L.raw.bn2_var, w->bn2_var, 16);
    memcpy(L.raw.last_w, w->last_w, sizeof(L.raw.last_w)); memcpy(L.raw.last_b, w->last_b, sizeof(L.raw.last_b));
    memcpy(L.raw.last_logs, w->last_logs, 16);
    L.raw.rescaling_scale = w->rescaling_scale;
    L.raw.bn_eps = w->bn_eps;
}

// Re-fold a coupling with explicit BatchNorm statistics: bn = {mean1[4], var1[4], mean2[4], var2[4]}.
int refold_with_stats(const Layer& L, const float* bn, NfCouplingP* out) {
    nf_coupling_weights w;
    w.l1_w = L.raw.l1_w; w.l1_b = L.raw.l1_b; w.bn1_mean = bn; w.bn1_var = bn + 4;
    w.l2_w = L.raw.l2_w; w.l2_b = L.raw.l2_b; w.bn2_mean = bn + 8; w.bn2_var = bn + 12;
    w.last_w = L.raw.last_w; w.last_b = L.raw.last_b; w.last_logs = L.raw.last_logs;
    w.rescaling_scale = L.raw.rescaling_scale; w.bn_eps = L.raw.bn_eps;
    return fold_coupling(&w, out);
}

void identity4(float m[4][4]) {
    for (int o = 0; o < 4; ++o)
        for (int i = 0; i < 4; ++i) m[o][i] = o == i ? 1.f : 0.f;
}

int set_scale_table(Layer& L, const float* table, int n_rows) {
    if (!table || n_rows < 1 || n_rows > NF_MAX_ROWS)
        return fail(NF_ERR_INVALID, "scale table: need 1..%d rows", NF_MAX_ROWS);
    L.n_rows = n_rows;
    memset(L.table, 0, sizeof(L.table));
    for (int r = 0; r < n_rows; ++r) {
        const double p0 = table[2 * r], p1 = table[2 * r + 1];
        if (L.scale_kind == NF_SCALE_SDN) {
            L.table[r][0] = (float)p0;
            L.table[r][1] = (float)p1;
        } else {
            // scale = g: inverse x /= g, log-det -sum log g (AffineCouplingGainEx4.py:115-124);
            // quirk variants return -log g without the sum over the 4096 dimensions.
            L.table[r][0] = (float)p0;
            L.table[r][1] = (float)(1.0 / p0);
            L.table[r][2] = (float)(-(L.full_sum ? (double)NF_DIMS : 1.0) * log(p0));
        }
    }
    return NF_OK;
}

// Build the kernel program for the bijectors [first, last): a conv1x1 / permutation that is
// immediately followed (data->latent order) by a coupling inside the range is fused into it.
int build_program(const nf_model* m, int first, int last, NfModelParams* mp, float* ldj_const,
                  int bn_layer = -1, const float* bn_stats = nullptr) {
    memset(mp, 0, sizeof(*mp));
    double ldj = 0.0;
    int n_cp = 0, n_mix = 0, n_sc = 0, n_ops = 0, n_rows = 0;
    for (int l = first; l < last; ++l) {
        const Layer& L = m->layers[l];
        if (n_ops >= NF_MAX_LAYERS) return fail(NF_ERR_UNSUPPORTED, "more than %d kernel ops", NF_MAX_LAYERS);
        if (L.kind == L_CONV1X1 || L.kind == L_PERMUTE) {
            ldj += L.log_abs_det * (double)NF_PIXELS;   // layers.py:129-130
            const bool fuse = (l + 1 < last) && m->layers[l + 1].kind == L_COUPLING;
            if (fuse) continue;   // picked up by the coupling below
            if (n_mix >= NF_MAX_MIX) return fail(NF_ERR_UNSUPPORTED, "more than %d stand-alone 1x1 layers", NF_MAX_MIX);
            memcpy(mp->mix[n_mix].a, L.a, sizeof(L.a));
            memcpy(mp->mix[n_mix].ainv, L.ainv, sizeof(L.ainv));
            mp->op[n_ops] = NF_KOP_MIX;
            mp->slot[n_ops++] = n_mix++;
        } else if (L.kind == L_COUPLING) {
            if (n_cp >= NF_MAX_COUPLINGS) return fail(NF_ERR_UNSUPPORTED, "more than %d couplings", NF_MAX_COUPLINGS);
            NfCouplingP& C = mp->cp[n_cp];
            C = L.cp;
            if (l == bn_layer) {   // batch-statistics BatchNorm for this coupling
                int rcf = refold_with_stats(L, bn_stats, &C);
                if (rcf) return rcf;
            }
            const bool fused = l > first && (m->layers[l - 1].kind == L_CONV1X1 || m->layers[l - 1].kind == L_PERMUTE);
            if (fused) {
                memcpy(C.a, m->layers[l - 1].a, sizeof(C.a));
                memcpy(C.ainv, m->layers[l - 1].ainv, sizeof(C.ainv));
                C.has_mix = 1;
            } else {
                identity4(C.a);
                identity4(C.ainv);
                C.has_mix = 0;
            }
            mp->op[n_ops] = NF_KOP_COUPLING;
            mp->slot[n_ops++] = n_cp++;
        } else if (L.kind == L_SCALE) {
            if (n_sc >= NF_MAX_SCALE) return fail(NF_ERR_UNSUPPORTED, "more than %d scale layers", NF_MAX_SCALE);
            memcpy(mp->sc[n_sc].t, L.table, sizeof(L.table));
            if (L.n_rows > n_rows) n_rows = L.n_rows;
            mp->op[n_ops] = L.scale_kind == NF_SCALE_SDN ? NF_KOP_SDN : NF_KOP_GAIN;
            mp->slot[n_ops++] = n_sc++;
        }
    }
    mp->n_layers = n_ops;
    mp->n_rows = n_rows;
    *ldj_const = (float)ldj;
    return NF_OK;
}

bool range_has_sdn(const nf_model* m, int first, int last) {      // ... or anything else that reads the clean patch
    for (int l = first; l < last; ++l) {
        if (m->layers[l].kind == L_SCALE && m->layers[l].scale_kind == NF_SCALE_SDN) return true;
        if (m->layers[l].kind == L_COUPLING && m->layers[l].cmode != 0) return true;
    }
    return false;
}

// ---- wide nets: fold one coupling into its blob block (NfWideLayout), double precision --------------------------
// bn: explicit BatchNorm statistics [mean1 W][var1 W][mean2 W][var2 W] or null -> the stored moving statistics
int fold_wide(const Layer& L, int W, const float* bn, const Layer* mixl, float* out) {
    const int CIN = L.cin, COUT = L.cout;
    const NfWideLayoutG lay(W, CIN, COUT);
    const WideRawView r(L.wraw.data(), W, CIN, COUT);
    const float *m1 = bn ? bn : r.bn1_mean, *v1 = bn ? bn + W : r.bn1_var, *m2 = bn ? bn + 2 * W : r.bn2_mean,
                *v2 = bn ? bn + 3 * W : r.bn2_var;
    const double eps = L.raw.bn_eps;
    const int oA = lay.A, oAINV = lay.AINV, oMETA = lay.META, oB3 = lay.B3, oB1 = lay.B1, oB2 = lay.b2(), oW1 = lay.w1(), oW2 = lay.w2(),
              oW3 = lay.w3();
    std::vector<double> s1(W), s2(W);
    for (int o = 0; o < W; ++o) {
        if (!((double)v1[o] + eps > 0.0) || !((double)v2[o] + eps > 0.0)) return fail(NF_ERR_INVALID, "batch-norm variance + eps must be positive");
        s1[o] = 1.0 / sqrt((double)v1[o] + eps);
        s2[o] = 1.0 / sqrt((double)v2[o] + eps);
    }
    memset(out, 0, (size_t)lay.size() * sizeof(float));
    double e[8];
    for (int o = 0; o < COUT; ++o) e[o] = exp(3.0 * (double)r.last_logs[o]);
    for (int o = 0; o < 4; ++o)
        for (int i = 0; i < 4; ++i) {
            out[oA + o * 4 + i] = mixl ? mixl->a[o][i] : (o == i ? 1.f : 0.f);
            out[oAINV + o * 4 + i] = mixl ? mixl->ainv[o][i] : (o == i ? 1.f : 0.f);
        }
    out[oMETA] = mixl ? 1.f : 0.f;
    out[oMETA + 1] = L.raw.rescaling_scale;
    out[oMETA + 2] = (float)L.cmode;
    out[oMETA + 3] = 0.f;
    for (int t = 0; t < 9; ++t)
        for (int o = 0; o < W; ++o)
            for (int i = 0; i < CIN; ++i) out[oW1 + (t * W + o) * CIN + i] = (float)((double)r.l1_w[(t * CIN + i) * W + o] * s1[o]);
    for (int o = 0; o < W; ++o) {
        out[oB1 + o] = (float)(((double)r.l1_b[o] - (double)m1[o]) * s1[o]);
        out[oB2 + o] = (float)(((double)r.l2_b[o] - (double)m2[o]) * s2[o]);
        for (int i = 0; i < W; ++i) out[oW2 + o * W + i] = (float)((double)r.l2_w[i * W + o] * s2[o]);
    }
    for (int t = 0; t < 9; ++t)
        for (int i = 0; i < W; ++i)
            for (int o = 0; o < COUT; ++o) out[oW3 + (t * W + i) * COUT + o] = (float)((double)r.last_w[(t * (W + 1) + i) * COUT + o] * e[o]);
    for (int rc = 0; rc < 3; ++rc)        // edge indicator -> bias table by (row class, column class), as fold_coupling
        for (int cc = 0; cc < 3; ++cc)
            for (int o = 0; o < COUT; ++o) {
                double acc = r.last_b[o];
                for (int dy = 0; dy < 3; ++dy)
                    for (int dx = 0; dx < 3; ++dx) {
                        const bool ring = (rc == 0 && dy == 0) || (rc == 2 && dy == 2) || (cc == 0 && dx == 0) || (cc == 2 && dx == 2);
                        if (ring) acc += (double)r.last_w[((dy * 3 + dx) * (W + 1) + W) * COUT + o];
                    }
                out[oB3 + (rc * 3 + cc) * COUT + o] = (float)(acc * e[o]);
            }
    return NF_OK;
}

// ---- tensor-core kernel: fold one coupling into its UMMA operand block (NfWideTcLayout), double precision, then bf16 (hi, lo)
uint16_t bf16_rn(float f) {            // round to nearest even
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);
    return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
float bf16_to_f(uint16_t b) {
    const uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
void bf16_split(double v, uint16_t* hi, uint16_t* lo) {
    *hi = bf16_rn((float)v);
    *lo = bf16_rn((float)(v - (double)bf16_to_f(*hi)));
}

int fold_wide_tc(const Layer& L, int W, const float* bn, const Layer* mixl, float* out) {
    using T = NfWideTcLayout;
    const WideRawView r(L.wraw.data(), W);
    const float *m1 = bn ? bn : r.bn1_mean, *v1 = bn ? bn + W : r.bn1_var, *m2 = bn ? bn + 2 * W : r.bn2_mean,
                *v2 = bn ? bn + 3 * W : r.bn2_var;
    const double eps = L.raw.bn_eps;
    std::vector<double> s1(W), s2(W);
    for (int o = 0; o < W; ++o) {
        if (!((double)v1[o] + eps > 0.0) || !((double)v2[o] + eps > 0.0)) return fail(NF_ERR_INVALID, "batch-norm variance + eps must be positive");
        s1[o] = 1.0 / sqrt((double)v1[o] + eps);
        s2[o] = 1.0 / sqrt((double)v2[o] + eps);
    }
    double e[4];
    for (int o = 0; o < 4; ++o) e[o] = exp(3.0 * (double)r.last_logs[o]);
    memset(out, 0, (size_t)T::block_bytes(W));
    for (int o = 0; o < 4; ++o)
        for (int i = 0; i < 4; ++i) {
            out[T::H_A + o * 4 + i] = mixl ? mixl->a[o][i] : (o == i ? 1.f : 0.f);
            out[T::H_AINV + o * 4 + i] = mixl ? mixl->ainv[o][i] : (o == i ? 1.f : 0.f);
        }
    out[T::H_META] = mixl ? 1.f : 0.f;
    out[T::H_META + 1] = L.raw.rescaling_scale;
    for (int rc = 0; rc < 3; ++rc)        // edge indicator -> bias table by (row class, column class), as fold_coupling
        for (int cc = 0; cc < 3; ++cc)
            for (int o = 0; o < 4; ++o) {
                double acc = r.last_b[o];
                for (int dy = 0; dy < 3; ++dy)
                    for (int dx = 0; dx < 3; ++dx) {
                        const bool ring = (rc == 0 && dy == 0) || (rc == 2 && dy == 2) || (cc == 0 && dx == 0) || (cc == 2 && dx == 2);
                        if (ring) acc += (double)r.last_w[((dy * 3 + dx) * (W + 1) + W) * 4 + o];
                    }
                out[T::H_B3 + (rc * 3 + cc) * 4 + o] = (float)(acc * e[o]);
            }
    unsigned char* base = reinterpret_cast<unsigned char*>(out);
    auto at = [](unsigned char* mat, int N, int n, int k) { return reinterpret_cast<uint16_t*>(mat) + ((size_t)(k / 8) * N + n) * 8 + k % 8; };
    uint16_t hi, lo;
    // B1: K rows [W1_hi (tap, in) | W1_hi | W1_lo | b_hi | b_lo]
    // widths 256 / 512 (nf_wide_tcs.cu, streamed weights): the same matrices cut into the blocks the ring carries --
    // B1 by 64 output channels, B2 by (pass of 256 outputs, 64 input channels), the bias by pass; B3 is unchanged
    const bool streamed = W >= 256;
    const int NK = W / 64;
    unsigned char* B1 = base + T::off_b1();
    for (int o = 0; o < W; ++o) {
        unsigned char* m = streamed ? B1 + (size_t)(o / 64) * 8192 : B1;
        const int N = streamed ? 64 : W, n = streamed ? o % 64 : o;
        for (int t = 0; t < 9; ++t)
            for (int i = 0; i < 2; ++i) {
                bf16_split((double)r.l1_w[(t * 2 + i) * W + o] * s1[o], &hi, &lo);
                *at(m, N, n, t * 2 + i) = hi;
                *at(m, N, n, 18 + t * 2 + i) = hi;
                *at(m, N, n, 36 + t * 2 + i) = lo;
            }
        bf16_split(((double)r.l1_b[o] - (double)m1[o]) * s1[o], &hi, &lo);
        *at(m, N, n, 54) = hi;
        *at(m, N, n, 55) = lo;
    }
    // B2: N rows [W2_hi | W2_lo], BB2: bias rows 6, 7
    unsigned char *B2 = base + T::off_b2(W), *BB2 = base + T::off_bb2(W);
    for (int o = 0; o < W; ++o) {
        for (int i = 0; i < W; ++i) {
            bf16_split((double)r.l2_w[i * W + o] * s2[o], &hi, &lo);
            if (streamed) {
                unsigned char* m = B2 + ((size_t)(o / 256) * NK + i / 64) * 65536;
                *at(m, 512, o % 256, i % 64) = hi;
                *at(m, 512, 256 + o % 256, i % 64) = lo;
            } else {
                *at(B2, 2 * W, o, i) = hi;
                *at(B2, 2 * W, W + o, i) = lo;
            }
        }
        bf16_split(((double)r.l2_b[o] - (double)m2[o]) * s2[o], &hi, &lo);
        unsigned char* mb = streamed ? BB2 + (size_t)(o / 256) * 8192 : BB2;
        *at(mb, streamed ? 256 : W, streamed ? o % 256 : o, 6) = hi;
        *at(mb, streamed ? 256 : W, streamed ? o % 256 : o, 7) = lo;
    }
    // B3: N row dy*16 + dx*4 + o (hi), 48 + ... (lo)
    unsigned char* B3 = base + T::off_b3(W);
    for (int t = 0; t < 9; ++t)
        for (int i = 0; i < W; ++i)
            for (int o = 0; o < 4; ++o) {
                bf16_split((double)r.last_w[(t * (W + 1) + i) * 4 + o] * e[o], &hi, &lo);
                const int n = (t / 3) * 16 + (t % 3) * 4 + o;
                *at(B3, 96, n, i) = hi;
                *at(B3, 96, 48 + n, i) = lo;
            }
    return NF_OK;
}

// kernel program + parameter blob of the bijectors [first, last) of a wide model (same fusion rule as build_program)
int build_wide_program(const nf_model* m, int first, int last, NfWideProgram* wp, std::vector<float>* blob, float* ldj_const,
                       int bn_layer = -1, const float* bn_stats = nullptr, bool tc = false) {
    const int W = m->width;
    memset(wp, 0, sizeof(*wp));
    wp->width = W;
    blob->clear();
    double ldj = 0.0;
    int n_ops = 0;
    for (int l = first; l < last; ++l) {
        const Layer& L = m->layers[l];
        if (n_ops >= NF_MAX_LAYERS) return fail(NF_ERR_UNSUPPORTED, "more than %d kernel ops", NF_MAX_LAYERS);
        if (tc && L.kind == L_COUPLING) blob->resize((blob->size() + NF_TMA_ROW_FLOATS - 1) / NF_TMA_ROW_FLOATS * NF_TMA_ROW_FLOATS, 0.f);   // TMA rows
        const size_t off = blob->size();
        if (L.kind == L_CONV1X1 || L.kind == L_PERMUTE) {
            ldj += L.log_abs_det * (double)NF_PIXELS;
            if ((l + 1 < last) && m->layers[l + 1].kind == L_COUPLING) continue;
            blob->resize(off + NF_WIDE_MIX_FLOATS);
            memcpy(blob->data() + off, L.a, sizeof(L.a));
            memcpy(blob->data() + off + 16, L.ainv, sizeof(L.ainv));
            wp->op[n_ops] = NF_KOP_MIX;
        } else if (L.kind == L_COUPLING) {
            if (tc && L.cmode != 0) return fail(NF_ERR_UNSUPPORTED, "clean-image-conditioned couplings run on the CUDA-core kernel (widths 4 / 8 / 16 / 32)");
            if (L.cmode != 0) wp->flags |= NF_WIDE_FLAG_COND;
            blob->resize(off + (size_t)(tc ? nf_wide_tc_coupling_floats(W) : nf_wide_coupling_floats(W, L.cin, L.cout)));
            const bool fused = l > first && (m->layers[l - 1].kind == L_CONV1X1 || m->layers[l - 1].kind == L_PERMUTE);
            int rc = tc ? fold_wide_tc(L, W, l == bn_layer ? bn_stats : nullptr, fused ? &m->layers[l - 1] : nullptr, blob->data() + off)
                        : fold_wide(L, W, l == bn_layer ? bn_stats : nullptr, fused ? &m->layers[l - 1] : nullptr, blob->data() + off);
            if (rc) return rc;
            wp->op[n_ops] = NF_KOP_COUPLING;
        } else if (L.kind == L_SCALE) {
            blob->resize(off + NF_WIDE_SCALE_FLOATS);
            memcpy(blob->data() + off, L.table, sizeof(L.table));
            wp->op[n_ops] = L.scale_kind == NF_SCALE_SDN ? NF_KOP_SDN : NF_KOP_GAIN;
        } else continue;
        wp->off[n_ops++] = (int32_t)off;
    }
    wp->n_layers = n_ops;
    if (tc) blob->resize((blob->size() + NF_TMA_ROW_FLOATS - 1) / NF_TMA_ROW_FLOATS * NF_TMA_ROW_FLOATS, 0.f);
    wp->blob_floats = (int32_t)blob->size();
    *ldj_const = (float)ldj;
    return NF_OK;
}

int check_ready(const nf_model* m) {
    if (!m) return fail(NF_ERR_INVALID, "null model");
    if (!m->finalized) return fail(NF_ERR_STATE, "model not finalized (call nf_model_finalize)");
    return NF_OK;
}

int num_ctas_for(const nf_model* m) { return m->num_ctas > 0 ? m->num_ctas : m->sm_count; }

// Wide model: launch the bijectors [first, last).  The full chain with the stored statistics uses the handle's resident
// blob; partial ranges and batch-statistics re-folds upload their own blob, stream-ordered (cudaMallocAsync).
// Tensor-core kernel (nf_wide_tc.cu) where the width has one (probes of the batch-statistics mode included).
int launch_wide_range(const nf_model* m, int first, int last, bool inverse, NfChainArgs& a, int bn_layer, const float* bn,
                      cudaStream_t stream, bool prefer_fp32 = false) {
    cudaError_t e;
    // prefer_fp32 (the train step): the backward kernels recompute activations in fp32, so the forward that stores every op's
    // input runs on the fp32 CUDA-core kernel where the width has one
    const bool tc = !m->has_cond && nf::wide_tc_width_supported(m->width) && (m->use_tc_wide || !nf::wide_width_supported(m->width)) &&
                    !(prefer_fp32 && nf::wide_width_supported(m->width));
    if (!tc && !nf::wide_width_supported(m->width))
        return fail(NF_ERR_UNSUPPORTED, "clean-image-conditioned couplings are built for coupling-net widths 4 / 8 / 16 / 32, got %d", m->width);
    if (first == 0 && last == (int)m->layers.size() && bn_layer < 0) {
        // the launch is enqueued under the lock: nf_model_finalize swaps and retires the blob under the same lock
        std::lock_guard<std::mutex> lock(m->prog_mu);
        const NfWideProgram& wp = tc ? m->wide_tc_full : m->wide_full;
        a.ldj_const = m->full_ldj_const;
        a.first_layer = 0;
        a.last_layer = wp.n_layers;
        e = tc ? nf::launch_chain_wide_tc(wp, m->d_wide_tc, a, inverse, num_ctas_for(m), stream)
               : nf::launch_chain_wide(wp, m->d_wide_full, a, inverse, num_ctas_for(m), stream);
    } else {
        NfWideProgram wp;
        std::vector<float> blob;
        float ldj = 0.f;
        int rc;
        {
            std::lock_guard<std::mutex> lock(m->prog_mu);
            rc = build_wide_program(m, first, last, &wp, &blob, &ldj, bn_layer, bn, tc);
        }
        if (rc) return rc;
        a.first_layer = 0;
        a.last_layer = wp.n_layers;
        a.ldj_const = ldj;
        float* d = nullptr;
        NF_CUDA(cudaMallocAsync((void**)&d, blob.size() * sizeof(float), stream));
        NF_CUDA(cudaMemcpyAsync(d, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice, stream));   // pageable: staged before return
        e = tc ? nf::launch_chain_wide_tc(wp, d, a, inverse, num_ctas_for(m), stream)
               : nf::launch_chain_wide(wp, d, a, inverse, num_ctas_for(m), stream);
        NF_CUDA(cudaFreeAsync(d, stream));
    }
    if (e != cudaSuccess) return fail(NF_ERR_CUDA, "wide chain kernel launch: %s", cudaGetErrorString(e));
    return NF_OK;
}

int launch_range(const nf_model* m, int first, int last, bool inverse, NfChainArgs& a, cudaStream_t stream) {
    if (a.n < 0) return fail(NF_ERR_INVALID, "negative patch count");
    if (a.n == 0) return NF_OK;
    if (!a.y && range_has_sdn(m, first, last)) return fail(NF_ERR_INVALID, "clean patch y is required by an sdn layer / a clean-image-conditioned coupling");
    if (a.default_row < 0 || a.default_row >= NF_MAX_ROWS) return fail(NF_ERR_INVALID, "default_row out of range");
    cudaError_t e;
    if (m->width != 4 || m->has_cond) return launch_wide_range(m, first, last, inverse, a, -1, nullptr, stream);
    if (first == 0 && last == (int)m->layers.size()) {
        NfModelParams mp;   // snapshot under the lock: parameters travel by value with the launch
        {
            std::lock_guard<std::mutex> lock(m->prog_mu);
            mp = m->full;
            a.ldj_const = m->full_ldj_const;
        }
        a.first_layer = 0;
        a.last_layer = mp.n_layers;
        if (a.in && nf::program_is_scale_only(mp, 0, mp.n_layers))
            e = nf::launch_scale_stream(mp, a, inverse, m->sm_count, stream);   // HBM-bound streaming path
        else if ((m->use_tc == 4 || m->use_tc == 0 || (m->use_tc == 3 && inverse)) && nf::wino_program_supported(mp, a))
            e = nf::launch_chain_wino(mp, a, inverse, num_ctas_for(m), stream);     // all-fp32, vertical Winograd F(2,3) convolutions
        else if ((m->use_tc == 2 || (m->use_tc == 3 && !inverse)) && nf::hybrid_program_supported(mp, a))
            e = nf::launch_chain_hybrid(mp, a, inverse, num_ctas_for(m), stream);   // conv-3 on tcgen05, the rest fp32
        else if (m->use_tc == 1 && nf::tc_program_supported(mp, a))
            e = nf::launch_chain_tc(mp, a, inverse, num_ctas_for(m), stream);   // tcgen05 path
        else
            e = nf::launch_chain(mp, a, inverse, num_ctas_for(m), m->warps_per_cta, stream);
    } else {
        NfModelParams mp;
        float ldj = 0.f;
        int rc;
        {
            std::lock_guard<std::mutex> lock(m->prog_mu);
            rc = build_program(m, first, last, &mp, &ldj);
        }
        if (rc) return rc;
        a.first_layer = 0;
        a.last_layer = mp.n_layers;
        a.ldj_const = ldj;
        if (a.in && nf::program_is_scale_only(mp, 0, mp.n_layers))
            e = nf::launch_scale_stream(mp, a, inverse, m->sm_count, stream);
        else if ((m->use_tc == 4 || m->use_tc == 0 || (m->use_tc == 3 && inverse)) && nf::wino_program_supported(mp, a))
            e = nf::launch_chain_wino(mp, a, inverse, num_ctas_for(m), stream);
        else if ((m->use_tc == 2 || (m->use_tc == 3 && !inverse)) && nf::hybrid_program_supported(mp, a))
            e = nf::launch_chain_hybrid(mp, a, inverse, num_ctas_for(m), stream);
        else if (m->use_tc == 1 && nf::tc_program_supported(mp, a))
            e = nf::launch_chain_tc(mp, a, inverse, num_ctas_for(m), stream);
        else
            e = nf::launch_chain(mp, a, inverse, num_ctas_for(m), m->warps_per_cta, stream);
    }
    if (e != cudaSuccess) return fail(NF_ERR_CUDA, "chain kernel launch: %s", cudaGetErrorString(e));
    return NF_OK;
}

// bijectors [lo, hi) with the BatchNorm of coupling `bn_layer` re-folded on explicit statistics (batch-statistics mode)
int launch_custom(const nf_model* m, int lo, int hi, bool inverse, NfChainArgs& a, int bn_layer, const float* bn, cudaStream_t stream,
                  bool prefer_fp32 = false) {
    if (m->width != 4 || m->has_cond) return launch_wide_range(m, lo, hi, inverse, a, bn_layer, bn, stream, prefer_fp32);
    NfModelParams mp;
    float ldjc = 0.f;
    int rc;
    {
        std::lock_guard<std::mutex> lock(m->prog_mu);
        rc = build_program(m, lo, hi, &mp, &ldjc, bn_layer, bn);
    }
    if (rc) return rc;
    a.first_layer = 0;
    a.last_layer = mp.n_layers;
    a.ldj_const = ldjc;
    cudaError_t e = nf::launch_chain(mp, a, inverse, num_ctas_for(m), m->warps_per_cta, stream);
    if (e != cudaSuccess) return fail(NF_ERR_CUDA, "chain launch: %s", cudaGetErrorString(e));
    return NF_OK;
}

}  // namespace

int nf::set_error(int code, const char* what, const char* msg) { return fail(code, "%s: %s", what, msg); }

// =====================================================================================================
extern "C" {

int nf_abi_version(void) { return NF_ABI_VERSION; }
const char* nf_last_error(void) { return g_err; }

int nf_device_info(int* sm_count, int* max_smem_optin, int* cc_major, int* cc_minor) {
    int dev = 0;
    NF_CUDA(cudaGetDevice(&dev));
    int v = 0;
    NF_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    if (sm_count) *sm_count = v;
    NF_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (max_smem_optin) *max_smem_optin = v;
    NF_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_major) *cc_major = v;
    NF_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
    if (cc_minor) *cc_minor = v;
    return NF_OK;
}

int nf_model_create(int height, int width, int channels, int net_width, nf_model** out) {
    if (!out) return fail(NF_ERR_INVALID, "out is null");
    *out = nullptr;
    if (height != NF_PATCH_H || width != NF_PATCH_W || channels != NF_PATCH_C)
        return fail(NF_ERR_UNSUPPORTED, "kernels are built for %dx%dx%d patches, got %dx%dx%d", NF_PATCH_H, NF_PATCH_W,
                    NF_PATCH_C, height, width, channels);
    if (net_width != 4 && !nf::wide_width_supported(net_width) && !nf::wide_tc_width_supported(net_width))
        return fail(NF_ERR_UNSUPPORTED, "kernels are built for coupling-net width 4 (fused warp-per-patch kernel), 8 / 16 / 32 "
                                        "(CTA-per-patch kernel) and 32 / 64 / 128 / 256 / 512 (tensor-core kernels), got %d", net_width);
    nf_model* m = new (std::nothrow) nf_model();
    if (!m) return fail(NF_ERR_INVALID, "out of host memory");
    m->width = net_width;
    *out = m;
    return NF_OK;
}

int nf_model_destroy(nf_model* m) {
    if (!m) return NF_OK;
    for (nf_model::HostPipe* hp : m->pipes_free) {
        for (auto& s : hp->st) {
            if (s.x) cudaFree(s.x);
            if (s.y) cudaFree(s.y);
            if (s.z) cudaFree(s.z);
            if (s.nll) cudaFree(s.nll);
            if (s.sdz) cudaFree(s.sdz);
            if (s.rows) cudaFree(s.rows);
            if (s.done) cudaEventDestroy(s.done);
            if (s.stream) cudaStreamDestroy(s.stream);
        }
        if (hp->h_tmp) cudaFreeHost(hp->h_tmp);
        delete hp;
    }
    if (m->d_wide_full) cudaFree(m->d_wide_full);
    if (m->d_wide_tc) cudaFree(m->d_wide_tc);
    for (auto& b : m->bs_free) cudaFree(b.first);
    delete m;
    return NF_OK;
}

static int fill_conv1x1(Layer& L, const float* A, const float* A_inv, float log_abs_det) {
    if (!A || !A_inv) return fail(NF_ERR_INVALID, "A / A_inv is null");
    L.kind = L_CONV1X1;
    for (int o = 0; o < 4; ++o)
        for (int i = 0; i < 4; ++i) {
            L.a[o][i] = A[i * 4 + o];          // inverse: x_out[o] = sum_i y[i] * A[i][o]   (layers.py:118-119)
            L.ainv[o][i] = A_inv[i * 4 + o];   // forward: uses A_inv                         (layers.py:113-114)
        }
    L.log_abs_det = log_abs_det;
    return NF_OK;
}

int nf_model_add_conv1x1(nf_model* m, const float* A, const float* A_inv, float log_abs_det) {
    if (!m) return fail(NF_ERR_INVALID, "null model");
    Layer L;
    int rc = fill_conv1x1(L, A, A_inv, log_abs_det);
    if (rc) return rc;
    m->layers.push_back(L);
    m->finalized = false;
    return NF_OK;
}

int nf_model_add_permute(nf_model* m, const int32_t* perm) {
    if (!m || !perm) return fail(NF_ERR_INVALID, "null argument");
    Layer L;
    L.kind = L_PERMUTE;
    bool seen[4] = {false, false, false, false};
    for (int i = 0; i < 4; ++i) {
        if (perm[i] < 0 || perm[i] > 3 || seen[perm[i]]) return fail(NF_ERR_INVALID, "perm is not a permutation of 0..3");
        seen[perm[i]] = true;
    }
    // forward y[i] = x[perm[i]]  -> ainv[o=i][in=perm[i]] = 1 ; inverse x[perm[i]] = y[i] -> a[o=perm[i]][in=i] = 1
    memset(L.a, 0, sizeof(L.a));
    memset(L.ainv, 0, sizeof(L.ainv));
    for (int i = 0; i < 4; ++i) {
        L.ainv[i][perm[i]] = 1.f;
        L.a[perm[i]][i] = 1.f;
    }
    L.log_abs_det = 0.0;
    m->layers.push_back(L);
    m->finalized = false;
    return NF_OK;
}

static int keep_wide_raw(Layer& L, int W, const nf_coupling_weights* w) {
    if (!w || !w->l1_w || !w->l1_b || !w->bn1_mean || !w->bn1_var || !w->l2_w || !w->l2_b || !w->bn2_mean ||
        !w->bn2_var || !w->last_w || !w->last_b || !w->last_logs)
        return fail(NF_ERR_INVALID, "nf_coupling_weights: null pointer");
    for (int o = 0; o < W; ++o)
        if (!(w->bn1_var[o] + w->bn_eps > 0.f) || !(w->bn2_var[o] + w->bn_eps > 0.f))
            return fail(NF_ERR_INVALID, "batch-norm variance + eps must be positive");
    const int cin = L.cin, cout = L.cout;
    L.wraw.resize(WideRawView::floats(W, cin, cout));
    float* p = L.wraw.data();
    auto put = [&](const float* src, size_t n) { memcpy(p, src, n * sizeof(float)); p += n; };
    put(w->l1_w, 9 * (size_t)cin * W); put(w->l1_b, W); put(w->bn1_mean, W); put(w->bn1_var, W);
    put(w->l2_w, (size_t)W * W); put(w->l2_b, W); put(w->bn2_mean, W); put(w->bn2_var, W);
    put(w->last_w, 9 * (size_t)cout * (W + 1)); put(w->last_b, cout); put(w->last_logs, cout);
    L.raw.rescaling_scale = w->rescaling_scale;
    L.raw.bn_eps = w->bn_eps;
    return NF_OK;
}

static int set_coupling_mode(Layer& L, int mode) {
    if (mode != NF_COUPLING_MODE_X && mode != NF_COUPLING_MODE_XY && mode != NF_COUPLING_MODE_Y) return fail(NF_ERR_INVALID, "unknown coupling mode %d", mode);
    L.cmode = mode;
    L.cin = mode == NF_COUPLING_MODE_X ? 2 : (mode == NF_COUPLING_MODE_XY ? 6 : 4);
    L.cout = mode == NF_COUPLING_MODE_Y ? 8 : 4;
    return NF_OK;
}

int nf_model_add_cond_coupling(nf_model* m, int mode, const nf_coupling_weights* w) {
    if (!m) return fail(NF_ERR_INVALID, "null model");
    if (mode == NF_COUPLING_MODE_X) return nf_model_add_affine_coupling(m, w);
    if (!nf::wide_width_supported(m->width))
        return fail(NF_ERR_UNSUPPORTED, "clean-image-conditioned couplings are built for coupling-net widths 4 / 8 / 16 / 32, got %d", m->width);
    Layer L;
    L.kind = L_COUPLING;
    int rc = set_coupling_mode(L, mode);
    if (!rc) rc = keep_wide_raw(L, m->width, w);
    if (rc) return rc;
    m->layers.push_back(L);
    m->has_cond = true;
    m->finalized = false;
    return NF_OK;
}

int nf_model_set_cond_coupling(nf_model* m, int layer, const nf_coupling_weights* w) {
    if (!m || layer < 0 || layer >= (int)m->layers.size() || m->layers[layer].kind != L_COUPLING || m->layers[layer].cmode == 0)
        return fail(NF_ERR_INVALID, "layer %d is not a clean-image-conditioned coupling", layer);
    int rc;
    {
        std::lock_guard<std::mutex> lock(m->prog_mu);
        rc = keep_wide_raw(m->layers[layer], m->width, w);
    }
    return rc ? rc : ((m->finalized && !m->defer_finalize) ? nf_model_finalize(m) : NF_OK);
}

int nf_model_add_affine_coupling(nf_model* m, const nf_coupling_weights* w) {
    if (!m) return fail(NF_ERR_INVALID, "null model");
    Layer L;
    L.kind = L_COUPLING;
    int rc;
    if (m->width != 4) rc = keep_wide_raw(L, m->width, w);
    else {
        rc = fold_coupling(w, &L.cp);
        if (!rc) keep_raw(L, w);
        if (!rc) rc = keep_wide_raw(L, 4, w);     // a model with clean-image-conditioned couplings runs on the generic kernel
    }
    if (rc) return rc;
    m->layers.push_back(L);
    m->finalized = false;
    return NF_OK;
}

int nf_model_add_scale(nf_model* m, int kind, int logdet_full_sum, const float* table, int n_rows) {
    if (!m) return fail(NF_ERR_INVALID, "null model");
    if (kind != NF_SCALE_SDN && kind != NF_SCALE_GAIN) return fail(NF_ERR_INVALID, "unknown scale kind %d", kind);
    Layer L;
    L.kind = L_SCALE;
    L.scale_kind = kind;
    L.full_sum = logdet_full_sum ? 1 : 0;
    int rc = set_scale_table(L, table, n_rows);
    if (rc) return rc;
    m->layers.push_back(L);
    m->finalized = false;
    return NF_OK;
}

int nf_model_finalize(nf_model* m) {
    if (!m) return fail(NF_ERR_INVALID, "null model");
    if (m->layers.empty()) return fail(NF_ERR_STATE, "model has no layers");
    int rc = NF_OK;
    if (m->width != 4 || m->has_cond) {   // wide / generic net: fold the whole chain and upload it to a FRESH device blob, then swap and retire the old one
        // (kernels already enqueued keep reading the old blob; cudaFree waits for them)
        for (int tc = 0; tc < 2; ++tc) {
            if (tc ? (m->has_cond || !nf::wide_tc_width_supported(m->width)) : !nf::wide_width_supported(m->width)) continue;
            std::vector<float> blob;
            NfWideProgram wp;
            float ldj = 0.f;
            {
                std::lock_guard<std::mutex> lock(m->prog_mu);
                rc = build_wide_program(m, 0, (int)m->layers.size(), &wp, &blob, &ldj, -1, nullptr, tc != 0);
            }
            if (rc) return rc;
            float* fresh = nullptr;
            NF_CUDA(cudaMalloc((void**)&fresh, blob.size() * sizeof(float)));
            cudaError_t ce = cudaMemcpy(fresh, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice);
            if (ce != cudaSuccess) { cudaFree(fresh); return fail(NF_ERR_CUDA, "wide blob upload: %s", cudaGetErrorString(ce)); }
            float* old = nullptr;
            {
                std::lock_guard<std::mutex> lock(m->prog_mu);
                if (tc) { old = m->d_wide_tc; m->d_wide_tc = fresh; m->wide_tc_full = wp; }
                else { old = m->d_wide_full; m->d_wide_full = fresh; m->wide_full = wp; m->wide_full_floats = blob.size(); }
                m->full_ldj_const = ldj;
            }
            if (old) cudaFree(old);
        }
    } else {
        std::lock_guard<std::mutex> lock(m->prog_mu);
        rc = build_program(m, 0, (int)m->layers.size(), &m->full, &m->full_ldj_const);
    }
    if (rc) return rc;
    if (m->sm_count == 0) {
        int sms = 0;
        rc = nf_device_info(&sms, nullptr, nullptr, nullptr);
        if (rc) return rc;
        m->sm_count = sms;
    }
    m->finalized = true;
    return NF_OK;
}

int nf_model_num_layers(const nf_model* m) { return m ? (int)m->layers.size() : fail(NF_ERR_INVALID, "null model"); }

static int refinalize(nf_model* m) { return (m->finalized && !m->defer_finalize) ? nf_model_finalize(m) : NF_OK; }

int nf_model_begin_update(nf_model* m) {
    if (!m) return fail(NF_ERR_INVALID, "null model");
    m->defer_finalize = 1;
    return NF_OK;
}
int nf_model_end_update(nf_model* m) {
    if (!m) return fail(NF_ERR_INVALID, "null model");
    m->defer_finalize = 0;
    return m->finalized ? nf_model_finalize(m) : NF_OK;
}

int nf_model_set_conv1x1(nf_model* m, int layer, const float* A, const float* A_inv, float log_abs_det) {
    if (!m || layer < 0 || layer >= (int)m->layers.size() || m->layers[layer].kind != L_CONV1X1)
        return fail(NF_ERR_INVALID, "layer %d is not a conv1x1", layer);
    int rc;
    {
        std::lock_guard<std::mutex> lock(m->prog_mu);
        rc = fill_conv1x1(m->layers[layer], A, A_inv, log_abs_det);
    }
    return rc ? rc : refinalize(m);
}

int nf_model_set_affine_coupling(nf_model* m, int layer, const nf_coupling_weights* w) {
    if (!m || layer < 0 || layer >= (int)m->layers.size() || m->layers[layer].kind != L_COUPLING)
        return fail(NF_ERR_INVALID, "layer %d is not an affine coupling", layer);
    int rc;
    {
        std::lock_guard<std::mutex> lock(m->prog_mu);
        if (m->layers[layer].cmode != 0) rc = fail(NF_ERR_INVALID, "layer %d is a clean-image-conditioned coupling: use nf_model_set_cond_coupling", layer);
        else if (m->width != 4) rc = keep_wide_raw(m->layers[layer], m->width, w);
        else {
            rc = fold_coupling(w, &m->layers[layer].cp);
            if (!rc) keep_raw(m->layers[layer], w);
            if (!rc) rc = keep_wide_raw(m->layers[layer], 4, w);
        }
    }
    return rc ? rc : refinalize(m);
}

int nf_model_set_scale(nf_model* m, int layer, const float* table, int n_rows) {
    if (!m || layer < 0 || layer >= (int)m->layers.size() || m->layers[layer].kind != L_SCALE)
        return fail(NF_ERR_INVALID, "layer %d is not a scale layer", layer);
    int rc;
    {
        std::lock_guard<std::mutex> lock(m->prog_mu);
        rc = set_scale_table(m->layers[layer], table, n_rows);
    }
    return rc ? rc : refinalize(m);
}

int nf_model_set_launch(nf_model* m, int warps_per_cta, int num_ctas) {
    if (!m) return fail(NF_ERR_INVALID, "null model");
    if (warps_per_cta < 1 || warps_per_cta > NF_MAX_WARPS_PER_CTA)
        return fail(NF_ERR_INVALID, "warps_per_cta must be in [1, %d]", NF_MAX_WARPS_PER_CTA);
    if (num_ctas < 0) return fail(NF_ERR_INVALID, "num_ctas must be >= 0");
    m->warps_per_cta = warps_per_cta;
    m->num_ctas = num_ctas;
    return NF_OK;
}

int nf_model_set_tensor_cores(nf_model* m, int enable) {
    if (!m) return fail(NF_ERR_INVALID, "null model");
    if (m->width == 4) m->use_tc = (enable >= 0 && enable <= 5) ? enable : 0;
    else {
        if (!enable && !nf::wide_width_supported(m->width))
            return fail(NF_ERR_UNSUPPORTED, "width %d has no CUDA-core kernel: the tensor-core kernel cannot be switched off", m->width);
        m->use_tc_wide = enable ? 1 : 0;
    }
    return NF_OK;
}

int nf_model_set_bs_small(nf_model* m, int enable) {
    if (!m) return fail(NF_ERR_INVALID, "null model");
    m->bs_small = enable ? 1 : 0;
    return NF_OK;
}

// ---- hot path -------------------------------------------------------------------------------------
int nf_log_prob(const nf_model* m, const float* x, const float* y, const int32_t* rows, int32_t default_row, int64_t n,
                float* nll, float* sdz, float* z, void* stream) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (n == 0) return NF_OK;
    if (!x || !nll) return fail(NF_ERR_INVALID, "x and nll are required");
    NfChainArgs a = {};
    a.in = x; a.y = y; a.rows = rows; a.out = z; a.nll = nll; a.sdz = sdz; a.n = n; a.default_row = default_row; a.temp = 1.f;
    return launch_range(m, 0, (int)m->layers.size(), true, a, (cudaStream_t)stream);
}

int nf_inverse(const nf_model* m, const float* x, const float* y, const int32_t* rows, int32_t default_row, int64_t n,
               float* z, float* logdet, void* stream) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (n == 0) return NF_OK;
    if (!x || !z) return fail(NF_ERR_INVALID, "x and z are required");
    NfChainArgs a = {};
    a.in = x; a.y = y; a.rows = rows; a.out = z; a.logdet = logdet; a.n = n; a.default_row = default_row; a.temp = 1.f;
    return launch_range(m, 0, (int)m->layers.size(), true, a, (cudaStream_t)stream);
}

int nf_forward(const nf_model* m, const float* z, const float* y, const int32_t* rows, int32_t default_row, int64_t n,
               float* x, float* logdet, void* stream) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (n == 0) return NF_OK;
    if (!z || !x) return fail(NF_ERR_INVALID, "z and x are required");
    NfChainArgs a = {};
    a.in = z; a.y = y; a.rows = rows; a.out = x; a.logdet = logdet; a.n = n; a.default_row = default_row; a.temp = 1.f;
    return launch_range(m, 0, (int)m->layers.size(), false, a, (cudaStream_t)stream);
}

int nf_sample(const nf_model* m, const float* y, const int32_t* rows, int32_t default_row, int64_t n, float temp,
              const float* eps, uint64_t seed, uint64_t offset, uint64_t patch_base, float* x, void* stream) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (n == 0) return NF_OK;
    if (!x) return fail(NF_ERR_INVALID, "x is required");
    NfChainArgs a = {};
    a.in = eps; a.y = y; a.rows = rows; a.out = x; a.n = n; a.default_row = default_row; a.temp = temp;
    a.seed = seed; a.offset = offset; a.patch_base = patch_base;
    return launch_range(m, 0, (int)m->layers.size(), false, a, (cudaStream_t)stream);
}

int nf_run_layers(const nf_model* m, int first, int last, int direction, const float* in, const float* y,
                  const int32_t* rows, int32_t default_row, int64_t n, float* out, float* logdet, void* stream) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (first < 0 || last > (int)m->layers.size() || first >= last) return fail(NF_ERR_INVALID, "bad layer range [%d,%d)", first, last);
    if (n == 0) return NF_OK;
    if (!in || !out) return fail(NF_ERR_INVALID, "in and out are required");
    if (direction != 0 && direction != 1) return fail(NF_ERR_INVALID, "direction must be 0 (inverse) or 1 (forward)");
    NfChainArgs a = {};
    a.in = in; a.y = y; a.rows = rows; a.out = out; a.logdet = logdet; a.n = n; a.default_row = default_row; a.temp = 1.f;
    return launch_range(m, first, last, direction == 0, a, (cudaStream_t)stream);
}

int nf_reduce_sums(const float* nll, const float* sdz, int64_t n, double* sums, void* stream) {
    if (!sums || n < 0) return fail(NF_ERR_INVALID, "sums is null or n < 0");
    cudaError_t e = nf::launch_reduce(nll, sdz, n, sums, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(NF_ERR_CUDA, "reduce kernel launch: %s", cudaGetErrorString(e));
    return NF_OK;
}

static int squeeze_common(const float* x, int64_t n, int H, int W, int C, int factor, int type, float* out, void* stream, int inv) {
    if (!x || !out || n < 0) return fail(NF_ERR_INVALID, "null pointer or n < 0");
    if (factor < 1 || H < 1 || W < 1 || C < 1 || H % factor || W % factor) return fail(NF_ERR_INVALID, "H and W must be multiples of factor");
    if (x == out && factor != 1) return fail(NF_ERR_INVALID, "squeeze cannot run in place");
    if (factor == 1) {   // identity (utils.py:32-33,65-66)
        if (x != out) NF_CUDA(cudaMemcpyAsync(out, x, (size_t)n * H * W * C * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return NF_OK;
    }
    cudaError_t e = nf::launch_squeeze(x, out, n, H, W, C, factor, type == 1 ? 1 : 0, inv, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(NF_ERR_CUDA, "squeeze kernel launch: %s", cudaGetErrorString(e));
    return NF_OK;
}
int nf_squeeze2d(const float* x, int64_t n, int H, int W, int C, int factor, int squeeze_type, float* out, void* stream) {
    return squeeze_common(x, n, H, W, C, factor, squeeze_type, out, stream, 0);
}
int nf_unsqueeze2d(const float* x, int64_t n, int H, int W, int C, int factor, int squeeze_type, float* out, void* stream) {
    return squeeze_common(x, n, H, W, C, factor, squeeze_type, out, stream, 1);
}

// ---- evaluation metrics next to the path ---------------------------------------------------------------
static int cached_sm_count() {
    static int sms_dev[NF_MAX_DEVICES] = {};   // per device
    int& sms = sms_dev[nf::device_slot()];
    if (!sms) { int v = 0; if (nf_device_info(&v, nullptr, nullptr, nullptr) == NF_OK) sms = v; }
    return sms ? sms : 148;
}

int nf_baseline_nll(const float* x, const float* y, float nlf0, float nlf1, float var_gauss, int64_t n, float* nll_gauss,
                    float* nll_sdn, void* stream) {
    if (n == 0) return NF_OK;
    if (n < 0 || !x || !y || (!nll_gauss && !nll_sdn)) return fail(NF_ERR_INVALID, "x, y and an output are required");
    if (!(var_gauss > 0.f)) return fail(NF_ERR_INVALID, "var_gauss must be positive");
    cudaError_t e = nf::launch_baseline_nll(x, y, nlf0, nlf1, var_gauss, n, nll_gauss, nll_sdn, cached_sm_count(), (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(NF_ERR_CUDA, "baseline kernel launch: %s", cudaGetErrorString(e));
    return NF_OK;
}

int nf_histogram(const float* data, int64_t count, const double* edges, int n_bins, unsigned long long* counts, void* stream) {
    if (count == 0) return NF_OK;
    if (count < 0 || !data || !edges || !counts) return fail(NF_ERR_INVALID, "null pointer or negative count");
    if (n_bins < 1 || n_bins > 4096) return fail(NF_ERR_INVALID, "n_bins must be in [1, 4096]");
    cudaError_t e = nf::launch_histogram(data, count, edges, n_bins, counts, cached_sm_count(), (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(NF_ERR_CUDA, "histogram kernel launch: %s", cudaGetErrorString(e));
    return NF_OK;
}

// ---- host-buffer pipeline ---------------------------------------------------------------------------
static const int64_t kChunk = 4096;   // patches per staged chunk: 64 MiB per tensor

namespace {
struct PipeLease {          // borrows a HostPipe for the duration of one _host call
    nf_model* m;
    nf_model::HostPipe* hp = nullptr;
    explicit PipeLease(nf_model* m_) : m(m_) {
        std::lock_guard<std::mutex> lock(m->pool_mu);
        if (!m->pipes_free.empty()) { hp = m->pipes_free.back(); m->pipes_free.pop_back(); }
        else { hp = new (std::nothrow) nf_model::HostPipe(); if (hp) ++m->pipes_total; }
    }
    ~PipeLease() {
        if (!hp) return;
        std::lock_guard<std::mutex> lock(m->pool_mu);
        m->pipes_free.push_back(hp);
    }
};

// slot k of the pipe, with buffers for `c` patches (grown on demand; streams / events created on first use)
int ensure_slot(nf_model::Staging& s, int64_t c, bool want_x, bool want_y, bool want_z, bool want_rows) {
    if (!s.stream) {
        NF_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        NF_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    const size_t pb = (size_t)NF_DIMS * sizeof(float);
    if (c > s.cap) {       // grow: everything the slot already holds is re-allocated at the new capacity
        NF_CUDA(cudaStreamSynchronize(s.stream));
        int64_t cap = 64;
        while (cap < c) cap *= 2;
        if (cap > kChunk) cap = kChunk;
        float** bufs[3] = {&s.x, &s.y, &s.z};
        for (auto b : bufs)
            if (*b) { cudaFree(*b); *b = nullptr; }
        if (s.nll) { cudaFree(s.nll); s.nll = nullptr; }
        if (s.sdz) { cudaFree(s.sdz); s.sdz = nullptr; }
        if (s.rows) { cudaFree(s.rows); s.rows = nullptr; }
        s.cap = cap;
    }
    if (want_x && !s.x) NF_CUDA(cudaMalloc(&s.x, s.cap * pb));
    if (want_y && !s.y) NF_CUDA(cudaMalloc(&s.y, s.cap * pb));
    if (want_z && !s.z) NF_CUDA(cudaMalloc(&s.z, s.cap * pb));
    if (!s.nll) NF_CUDA(cudaMalloc(&s.nll, s.cap * sizeof(float)));
    if (!s.sdz) NF_CUDA(cudaMalloc(&s.sdz, s.cap * sizeof(float)));
    if (want_rows && !s.rows) NF_CUDA(cudaMalloc(&s.rows, s.cap * sizeof(int32_t)));
    return NF_OK;
}
}  // namespace

int nf_log_prob_host(const nf_model* cm, const float* x_host, const float* y_host, const int32_t* rows_host,
                     int32_t default_row, int64_t n, float* nll_host, float* sdz_host, float* z_host, double* sums_host) {
    int rc = check_ready(cm);
    if (rc) return rc;
    if (!x_host || n < 0) return fail(NF_ERR_INVALID, "x_host is required");
    if (!nll_host && !sums_host) return fail(NF_ERR_INVALID, "nll_host or sums_host is required");
    nf_model* m = const_cast<nf_model*>(cm);
    PipeLease lease(m);            // this caller's own pipeline: concurrent callers do not serialise
    if (!lease.hp) return fail(NF_ERR_INVALID, "out of host memory");
    nf_model::HostPipe& hp = *lease.hp;
    const size_t pb = (size_t)NF_DIMS * sizeof(float);
    double tot[3] = {0.0, 0.0, 0.0};
    // Results the caller did not ask for but the sums need go to PINNED scratch: a device->host copy into pageable
    // memory blocks the host until the chunk's kernel has finished, which would serialise copies and compute.
    const size_t need = ((sums_host && !nll_host) ? (size_t)n : 0) + ((sums_host && !sdz_host) ? (size_t)n : 0);
    if (need > hp.h_tmp_floats) {
        if (hp.h_tmp) cudaFreeHost(hp.h_tmp);
        hp.h_tmp = nullptr;
        hp.h_tmp_floats = 0;
        NF_CUDA(cudaHostAlloc((void**)&hp.h_tmp, need * sizeof(float), cudaHostAllocDefault));
        hp.h_tmp_floats = need;
    }
    float* tmp_nll = (sums_host && !nll_host) ? hp.h_tmp : nullptr;
    float* tmp_sdz = (sums_host && !sdz_host) ? hp.h_tmp + (tmp_nll ? (size_t)n : 0) : nullptr;
    float* nll_dst = nll_host ? nll_host : tmp_nll;
    float* sdz_dst = sdz_host ? sdz_host : tmp_sdz;
    int64_t k = 0;
    for (int64_t off = 0; off < n; off += kChunk, ++k) {
        nf_model::Staging& s = hp.st[k % nf_model::kSlots];
        const int64_t c = (n - off < kChunk) ? n - off : kChunk;
        rc = ensure_slot(s, c, true, y_host != nullptr, z_host != nullptr, rows_host != nullptr);
        if (rc) return rc;
        NF_CUDA(cudaEventSynchronize(s.done));   // previous use of this slot has drained
        NF_CUDA(cudaMemcpyAsync(s.x, x_host + off * NF_DIMS, c * pb, cudaMemcpyHostToDevice, s.stream));
        if (y_host) NF_CUDA(cudaMemcpyAsync(s.y, y_host + off * NF_DIMS, c * pb, cudaMemcpyHostToDevice, s.stream));
        if (rows_host) NF_CUDA(cudaMemcpyAsync(s.rows, rows_host + off, c * sizeof(int32_t), cudaMemcpyHostToDevice, s.stream));
        rc = nf_log_prob(m, s.x, y_host ? s.y : nullptr, rows_host ? s.rows : nullptr, default_row, c, s.nll, s.sdz,
                         z_host ? s.z : nullptr, s.stream);
        if (rc) return rc;
        NF_CUDA(cudaMemcpyAsync(nll_dst + off, s.nll, c * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        if (sdz_dst) NF_CUDA(cudaMemcpyAsync(sdz_dst + off, s.sdz, c * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        if (z_host) NF_CUDA(cudaMemcpyAsync(z_host + off * NF_DIMS, s.z, c * pb, cudaMemcpyDeviceToHost, s.stream));
        NF_CUDA(cudaEventRecord(s.done, s.stream));
    }
    for (auto& s : hp.st)
        if (s.stream) NF_CUDA(cudaStreamSynchronize(s.stream));
    if (sums_host) {   // fixed-order fp64 accumulation on the host: identical for any chunking
        for (int64_t i = 0; i < n; ++i) {
            tot[0] += (double)nll_dst[i];
            if (sdz_dst) tot[1] += (double)sdz_dst[i];
        }
        tot[2] = (double)n;
        memcpy(sums_host, tot, sizeof(tot));
    }
    return NF_OK;
}

int nf_sample_host(const nf_model* cm, const float* y_host, const int32_t* rows_host, int32_t default_row, int64_t n,
                   float temp, const float* eps_host, uint64_t seed, uint64_t offset, float* x_host) {
    int rc = check_ready(cm);
    if (rc) return rc;
    if (!x_host || n < 0) return fail(NF_ERR_INVALID, "x_host is required");
    nf_model* m = const_cast<nf_model*>(cm);
    PipeLease lease(m);
    if (!lease.hp) return fail(NF_ERR_INVALID, "out of host memory");
    nf_model::HostPipe& hp = *lease.hp;
    const size_t pb = (size_t)NF_DIMS * sizeof(float);
    int64_t k = 0;
    for (int64_t off = 0; off < n; off += kChunk, ++k) {
        nf_model::Staging& s = hp.st[k % nf_model::kSlots];
        const int64_t c = (n - off < kChunk) ? n - off : kChunk;
        rc = ensure_slot(s, c, eps_host != nullptr, y_host != nullptr, true, rows_host != nullptr);
        if (rc) return rc;
        NF_CUDA(cudaEventSynchronize(s.done));
        if (y_host) NF_CUDA(cudaMemcpyAsync(s.y, y_host + off * NF_DIMS, c * pb, cudaMemcpyHostToDevice, s.stream));
        if (eps_host) NF_CUDA(cudaMemcpyAsync(s.x, eps_host + off * NF_DIMS, c * pb, cudaMemcpyHostToDevice, s.stream));
        if (rows_host) NF_CUDA(cudaMemcpyAsync(s.rows, rows_host + off, c * sizeof(int32_t), cudaMemcpyHostToDevice, s.stream));
        rc = nf_sample(m, y_host ? s.y : nullptr, rows_host ? s.rows : nullptr, default_row, c, temp,
                       eps_host ? s.x : nullptr, seed, offset, (uint64_t)off, s.z, s.stream);
        if (rc) return rc;
        NF_CUDA(cudaMemcpyAsync(x_host + off * NF_DIMS, s.z, c * pb, cudaMemcpyDeviceToHost, s.stream));
        NF_CUDA(cudaEventRecord(s.done, s.stream));
    }
    for (auto& s : hp.st)
        if (s.stream) NF_CUDA(cudaStreamSynchronize(s.stream));
    return NF_OK;
}

// ---- batch-statistics BatchNorm, small batches: the whole chain as ONE cooperative kernel ----------------
// (nf_trainer.cu: td_bs_chain_kernel).  Used by nf_chain_batch_stats when every patch can own a co-resident CTA (296 on a
// B200) and the chain is made of [mix +] coupling groups and scale layers at width 4; everything else takes the
// layer-by-layer path below.  The parameter image (raw TF-layout weights, matrices, tables) is rebuilt from the handle on
// every call -- 12 KB, and it makes the call independent of concurrent callers and of nf_model_set_* updates.
static int chain_batch_stats_small(const nf_model* m, const std::vector<std::pair<int, int>>& groups, bool inverse, const float* in,
                                   const float* y, const int32_t* rows, int32_t default_row, int64_t n, float temp, uint64_t seed,
                                   uint64_t offset, uint64_t patch_base, float* out, float* logdet, float* nll, float* sdz,
                                   float* batch_stats_host, cudaStream_t stream) {
    std::vector<nf::BsOp> ops;
    std::vector<float> vars, amat, tables;
    std::vector<int> cp_add_index;      // coupling (in execution order) -> its index in add order
    double ldj_const = 0.0;
    {
        std::lock_guard<std::mutex> lock(m->prog_mu);
        for (const auto& g : groups) {
            const Layer& L = m->layers[g.second - 1];
            nf::BsOp op = {};
            if (L.kind == L_COUPLING) {
                op.kind = 0;
                op.cidx = (int)cp_add_index.size();
                int idx = 0;
                for (int l = 0; l < g.second - 1; ++l) idx += m->layers[l].kind == L_COUPLING;
                cp_add_index.push_back(idx);
                auto push = [&](const float* p, int k) { const int o = (int)vars.size(); vars.insert(vars.end(), p, p + k); return o; };
                nf::TdCoupling& d = op.d;
                d.off_w1 = push(L.raw.l1_w, 72); d.off_b1 = push(L.raw.l1_b, 4);
                d.off_w2 = push(L.raw.l2_w, 16); d.off_b2 = push(L.raw.l2_b, 4);
                d.off_w3 = push(L.raw.last_w, 180); d.off_b3 = push(L.raw.last_b, 4);
                d.off_logs = push(L.raw.last_logs, 4); d.off_scale = push(&L.raw.rescaling_scale, 1);
                d.off_bn[0] = push(L.raw.bn1_mean, 4); d.off_bn[1] = push(L.raw.bn1_var, 4);
                d.off_bn[2] = push(L.raw.bn2_mean, 4); d.off_bn[3] = push(L.raw.bn2_var, 4);
                d.has_mix = g.second - g.first == 2;
                d.batch_stats = 1;
                d.bn_eps = L.raw.bn_eps;
                float A[16];
                for (int k = 0; k < 16; ++k) A[k] = (k >> 2) == (k & 3) ? 1.f : 0.f;
                if (d.has_mix) {
                    const Layer& M = m->layers[g.first];
                    for (int i = 0; i < 4; ++i)
                        for (int o = 0; o < 4; ++o) A[i * 4 + o] = inverse ? M.a[o][i] : M.ainv[o][i];     // [in][out]
                    ldj_const += (double)NF_PIXELS * M.log_abs_det;
                }
                amat.insert(amat.end(), A, A + 16);
            } else {
                op.kind = 1;
                op.sidx = (int)(tables.size() / (NF_MAX_ROWS * 2));
                op.is_sdn = L.scale_kind == NF_SCALE_SDN;
                op.full_sum = L.full_sum;
                for (int r = 0; r < NF_MAX_ROWS; ++r) { tables.push_back(L.table[r][0]); tables.push_back(op.is_sdn ? L.table[r][1] : 0.f); }
            }
            ops.push_back(op);
        }
    }
    const size_t n_cp = cp_add_index.size();
    // one device buffer: [stats 16 n_cp doubles][consts 1 double][ops][vars][amat][tables]
    auto up8 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t o_stats = 0, o_consts = o_stats + 16 * n_cp * sizeof(double), o_ops = up8(o_consts + sizeof(double)),
                 o_vars = up8(o_ops + ops.size() * sizeof(nf::BsOp)), o_amat = up8(o_vars + vars.size() * sizeof(float)),
                 o_tab = up8(o_amat + amat.size() * sizeof(float)), total = up8(o_tab + tables.size() * sizeof(float)) + 16;
    // host image of everything behind the statistics: ONE upload
    std::vector<unsigned char> img(total - o_consts, 0);
    auto put = [&](size_t off, const void* src, size_t bytes) { if (bytes) memcpy(img.data() + (off - o_consts), src, bytes); };
    put(o_consts, &ldj_const, sizeof(double));
    put(o_ops, ops.data(), ops.size() * sizeof(nf::BsOp));
    put(o_vars, vars.data(), vars.size() * sizeof(float));
    put(o_amat, amat.data(), amat.size() * sizeof(float));
    put(o_tab, tables.data(), tables.size() * sizeof(float));
    unsigned char* blob = nullptr;
    size_t blob_bytes = 0;
    {
        std::lock_guard<std::mutex> lock(m->bs_mu);
        if (!m->bs_free.empty()) { blob = m->bs_free.back().first; blob_bytes = m->bs_free.back().second; m->bs_free.pop_back(); }
    }
    if (blob && blob_bytes < total) { cudaFree(blob); blob = nullptr; }
    if (!blob) { NF_CUDA(cudaMalloc((void**)&blob, total)); blob_bytes = total; }
    int rc = NF_OK;
    std::vector<double> hst(16 * n_cp + 1);
    do {
        cudaError_t e;
        if ((e = cudaMemsetAsync(blob, 0, o_consts, stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(blob + o_consts, img.data(), img.size(), cudaMemcpyHostToDevice, stream)) != cudaSuccess) {   // pageable: staged before return
            rc = fail(NF_ERR_CUDA, "small-batch chain upload: %s", cudaGetErrorString(e));
            break;
        }
        nf::BsArgs a = {};
        a.ops = (const nf::BsOp*)(blob + o_ops); a.n_ops = (int)ops.size(); a.direction = inverse ? 0 : 1;
        a.vars = (const float*)(blob + o_vars); a.Amat = (const float*)(blob + o_amat); a.tables = (const float*)(blob + o_tab);
        a.stats = (double*)(blob + o_stats); a.consts = (const double*)(blob + o_consts);
        a.in = in; a.y = y; a.rows = rows; a.default_row = default_row; a.n = n; a.out = out;
        a.nll = inverse ? nll : nullptr; a.sdz = inverse ? sdz : nullptr; a.logdet = inverse ? logdet : nullptr;
        a.ld = inverse ? (logdet ? logdet : nll) : nullptr;      // running log-det (nll doubles as scratch), zeroed below
        a.temp = temp; a.seed = seed; a.offset = offset; a.patch_base = patch_base;
        if (a.ld && (e = cudaMemsetAsync(a.ld, 0, (size_t)n * sizeof(float), stream)) != cudaSuccess) { rc = fail(NF_ERR_CUDA, "memset: %s", cudaGetErrorString(e)); break; }
        if ((e = nf::launch_bs_small(a, nf::bs_small_capacity(cached_sm_count()), stream)) != cudaSuccess) {
            if (e == cudaErrorCooperativeLaunchTooLarge) { cudaGetLastError(); rc = 1; break; }   // e.g. SMs partitioned by MPS: layer by layer
            rc = fail(NF_ERR_CUDA, "small-batch chain launch: %s", cudaGetErrorString(e));
            break;
        }
        if ((e = cudaMemcpyAsync(hst.data(), blob + o_stats, 16 * n_cp * sizeof(double), cudaMemcpyDeviceToHost, stream)) != cudaSuccess) { rc = fail(NF_ERR_CUDA, "statistics read-back: %s", cudaGetErrorString(e)); break; }
    } while (0);
    cudaError_t es = cudaStreamSynchronize(stream);
    {
        std::lock_guard<std::mutex> lock(m->bs_mu);
        m->bs_free.push_back({blob, blob_bytes});
    }
    if (rc) return rc;
    if (es != cudaSuccess) return fail(NF_ERR_CUDA, "small-batch chain: %s", cudaGetErrorString(es));
    if (batch_stats_host) {
        const double cnt = (double)n * NF_PIXELS;
        for (size_t c = 0; c < n_cp; ++c) {
            float* dst = batch_stats_host + 16 * (size_t)cp_add_index[c];      // {mean1[4], var1[4], mean2[4], var2[4]}
            for (int st = 0; st < 2; ++st)
                for (int k = 0; k < 4; ++k) {
                    const double mean = hst[16 * c + 8 * st + k] / cnt;
                    double var = hst[16 * c + 8 * st + 4 + k] / cnt - mean * mean;     // population variance (tf.nn.moments)
                    if (var < 0.0) var = 0.0;
                    dst[8 * st + k] = (float)mean;
                    dst[8 * st + 4 + k] = (float)var;
                }
        }
    }
    return NF_OK;
}

// ---- batch-statistics BatchNorm: layer-by-layer execution ----------------------------------------------
// Reference: batch_norm(training=True) normalises every coupling-net activation with the statistics of the
// CURRENT batch (layers.py:388-398), which makes patches interdependent: 16 batch-wide reductions per
// pass.  Each coupling therefore runs as probe(conv-1 stats) -> probe(conv-2 stats) -> apply.
int nf_chain_batch_stats(const nf_model* m, int direction, const float* in, const float* y, const int32_t* rows,
                         int32_t default_row, int64_t n, float temp, uint64_t seed, uint64_t offset, uint64_t patch_base,
                         float* out, float* logdet, float* nll, float* sdz, double* stats_ws, float* batch_stats_host,
                         void* stream_) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (direction != 0 && direction != 1) return fail(NF_ERR_INVALID, "direction must be 0 (inverse) or 1 (forward)");
    if (n == 0) return NF_OK;
    if (n < 0 || !out || !stats_ws) return fail(NF_ERR_INVALID, "out and stats_ws are required");
    if (direction == 0 && !in) return fail(NF_ERR_INVALID, "in is required for the inverse direction");
    if (default_row < 0 || default_row >= NF_MAX_ROWS) return fail(NF_ERR_INVALID, "default_row out of range");
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool inverse = direction == 0;
    const int L = (int)m->layers.size(), W = m->width;
    // groups of bijectors that form one kernel op: [mix, coupling] pairs, or single layers
    std::vector<std::pair<int, int>> groups;
    for (int l = 0; l < L;) {
        const bool mix = m->layers[l].kind == L_CONV1X1 || m->layers[l].kind == L_PERMUTE;
        if (mix && l + 1 < L && m->layers[l + 1].kind == L_COUPLING) { groups.push_back({l, l + 2}); l += 2; }
        else { groups.push_back({l, l + 1}); l += 1; }
    }
    if (!inverse) std::reverse(groups.begin(), groups.end());
    // one cooperative kernel, no host round trip: one co-resident CTA per patch up to 296 patches (B200), the same kernel
    // walking its patches grid-stride up to kBsCoopMax; beyond that the probes of the warp-per-patch kernel win by more than the
    // 16 stream synchronisations cost (< 1 % of the call at that size)
    const int64_t kBsCoopMax = 4096;
    if (W == 4 && !m->has_cond && m->bs_small && !(direction == 1 && logdet) && nf::bs_small_capacity(cached_sm_count()) > 0 && n <= kBsCoopMax) {
        bool ok = true;       // [mix +] coupling groups and scale layers only
        for (const auto& g : groups) {
            const int k = m->layers[g.second - 1].kind;
            ok = ok && (k == L_COUPLING || (k == L_SCALE && g.second - g.first == 1));
            if (k == L_SCALE && m->layers[g.second - 1].scale_kind == NF_SCALE_SDN && !y) return fail(NF_ERR_INVALID, "clean patch y is required by an sdn layer");
        }
        if (ok) {
            rc = chain_batch_stats_small(m, groups, inverse, in, y, rows, default_row, n, temp, seed, offset, patch_base, out, logdet, nll,
                                         sdz, batch_stats_host, stream);
            if (rc != 1) return rc;     // 1: the cooperative grid does not fit right now -> the path below
        }
    }
    float* run_ld = logdet ? logdet : nll;   // running log-det (nll doubles as scratch until the last launch)
    const bool want_ld = run_ld != nullptr;
    const double cnt = (double)n * NF_PIXELS;
    std::vector<float> bn((size_t)4 * W);          // [mean1 W][var1 W][mean2 W][var2 W]
    std::vector<double> h((size_t)2 * W);
    bool first = true;
    for (size_t g = 0; g < groups.size(); ++g) {
        const int lo = groups[g].first, hi = groups[g].second;
        const int cl = m->layers[hi - 1].kind == L_COUPLING ? hi - 1 : -1;
        if (!y && range_has_sdn(m, lo, hi)) return fail(NF_ERR_INVALID, "clean patch y is required by an sdn layer");
        if (cl >= 0) {
            const float ident = 1.0f - m->layers[cl].raw.bn_eps;   // identity fold: 1/sqrt(var + eps) == 1
            for (int k = 0; k < W; ++k) { bn[k] = 0.f; bn[2 * W + k] = 0.f; bn[W + k] = ident; bn[3 * W + k] = ident; }
            for (int stage = 1; stage <= 2; ++stage) {
                NF_CUDA(cudaMemsetAsync(stats_ws, 0, 2 * (size_t)W * sizeof(double), stream));
                NfChainArgs a = {};
                a.in = first ? in : out; a.y = y; a.rows = rows; a.n = n; a.default_row = default_row;
                a.temp = first ? temp : 1.f; a.seed = seed; a.offset = offset; a.patch_base = patch_base;
                a.bn_stats = stats_ws; a.bn_stage = stage;
                rc = launch_custom(m, lo, hi, inverse, a, cl, bn.data(), stream);
                if (rc) return rc;
                NF_CUDA(cudaMemcpyAsync(h.data(), stats_ws, 2 * (size_t)W * sizeof(double), cudaMemcpyDeviceToHost, stream));
                NF_CUDA(cudaStreamSynchronize(stream));
                for (int k = 0; k < W; ++k) {
                    const double mean = h[k] / cnt;
                    double var = h[W + k] / cnt - mean * mean;   // population variance (tf.nn.moments)
                    if (var < 0.0) var = 0.0;
                    bn[(stage - 1) * 2 * W + k] = (float)mean;
                    bn[(stage - 1) * 2 * W + W + k] = (float)var;
                }
            }
            if (batch_stats_host) {
                int idx = 0;
                for (int l = 0; l < cl; ++l) idx += m->layers[l].kind == L_COUPLING;
                memcpy(batch_stats_host + (size_t)4 * W * idx, bn.data(), (size_t)4 * W * sizeof(float));
            }
        }
        const bool last = g + 1 == groups.size();
        NfChainArgs a = {};
        a.in = first ? in : out; a.y = y; a.rows = rows; a.out = out; a.n = n; a.default_row = default_row;
        a.temp = first ? temp : 1.f; a.seed = seed; a.offset = offset; a.patch_base = patch_base;
        a.logdet_in = (want_ld && !first) ? run_ld : nullptr;
        if (last) { a.logdet = logdet; a.nll = inverse ? nll : nullptr; a.sdz = inverse ? sdz : nullptr; }
        else a.logdet = want_ld ? run_ld : nullptr;
        rc = launch_custom(m, lo, hi, inverse, a, cl, cl >= 0 ? bn.data() : nullptr, stream);
        if (rc) return rc;
        first = false;
    }
    NF_CUDA(cudaStreamSynchronize(stream));
    return NF_OK;
}

// ---- training: loss + gradient of every trainable variable ---------------------------------------------
namespace {
struct Group { int lo, hi, cl; };   // bijectors [lo, hi) form one kernel op; cl = index of its coupling or -1

std::vector<Group> make_groups(const nf_model* m) {
    std::vector<Group> g;
    const int L = (int)m->layers.size();
    for (int l = 0; l < L;) {
        const bool mix = m->layers[l].kind == L_CONV1X1 || m->layers[l].kind == L_PERMUTE;
        if (mix && l + 1 < L && m->layers[l + 1].kind == L_COUPLING) { g.push_back({l, l + 2, l + 1}); l += 2; }
        else { g.push_back({l, l + 1, m->layers[l].kind == L_COUPLING ? l : -1}); l += 1; }
    }
    return g;
}

void fill_train_coupling(const nf_model* m, const Group& g, const float* bn, NfTrainCoupling* P) {
    const Layer& L = m->layers[g.cl];
    memset(P, 0, sizeof(*P));
    const bool fused = g.hi - g.lo == 2;
    P->has_mix = fused ? 1 : 0;
    for (int i = 0; i < 4; ++i)
        for (int o = 0; o < 4; ++o) P->A[i][o] = fused ? m->layers[g.lo].a[o][i] : (i == o ? 1.f : 0.f);   // a[o][i] = A[i][o]
    memcpy(P->w1, L.raw.l1_w, sizeof(P->w1));
    memcpy(P->w2, L.raw.l2_w, sizeof(P->w2));
    memcpy(P->w3, L.raw.last_w, sizeof(P->w3));
    for (int k = 0; k < 4; ++k) {
        P->b1[k] = L.raw.l1_b[k]; P->b2[k] = L.raw.l2_b[k]; P->b3[k] = L.raw.last_b[k]; P->logs[k] = L.raw.last_logs[k];
        P->m1[k] = bn[k];      P->is1[k] = (float)(1.0 / sqrt((double)bn[4 + k] + (double)L.raw.bn_eps));
        P->m2[k] = bn[8 + k];  P->is2[k] = (float)(1.0 / sqrt((double)bn[12 + k] + (double)L.raw.bn_eps));
    }
    P->scale = L.raw.rescaling_scale;
}

int64_t layer_grad_size(const Layer& L, int W = 4) {
    switch (L.kind) {
        case L_CONV1X1: return 16;
        case L_COUPLING: return W == 4 ? NF_G_HOST_COUPLING : nf_train_wide_host_coupling(W);
        case L_SCALE: return 2 * (int64_t)L.n_rows;
        default: return 0;
    }
}
}  // namespace

int nf_grad_layout(const nf_model* m, int64_t* offsets) {
    if (!m || !offsets) return fail(NF_ERR_INVALID, "null argument");
    int64_t off = 0;
    for (size_t l = 0; l < m->layers.size(); ++l) { offsets[l] = off; off += layer_grad_size(m->layers[l], m->width); }
    offsets[m->layers.size()] = off;
    return NF_OK;
}

int nf_train_workspace_floats(const nf_model* m, int64_t n, int64_t* n_floats) {
    if (!m || !n_floats || n < 0) return fail(NF_ERR_INVALID, "bad argument");
    const int64_t G = (int64_t)make_groups(m).size();
    // width 4: G op inputs + two gradient buffers + scratch + g_z'; wider nets: the scratch holds one float per hidden channel
    *n_floats = (G + 3 + (m->width == 4 ? 1 : m->width / 4)) * n * NF_DIMS + 2 * n;
    return NF_OK;
}

// Coupling nets wider than 4 (nf_train_wide.cu): same orchestration as below -- forward with batch-statistics BatchNorm keeping
// every op's input (probe + apply launches of the wide chain kernels), backward sweep with three passes per coupling and the
// BatchNorm sums read back in between.
static int loss_and_grad_wide(const nf_model* m, const float* x, const float* y, const int32_t* rows, int32_t default_row, int64_t n,
                              int batch_stats, float* workspace, double* dscratch, double* grads_host, float* batch_stats_host,
                              double* sums_host, cudaStream_t stream) {
    const int W = m->width;
    if (!nf::train_wide_width_supported(W)) return fail(NF_ERR_UNSUPPORTED, "the train-step kernels are built for coupling-net widths 4 / 8 / 16 / 32, got %d", W);
    int rc;
    const std::vector<Group> groups = make_groups(m);
    const int G = (int)groups.size();
    const int64_t S = n * NF_DIMS;
    auto slot = [&](int k) { return workspace + (int64_t)k * S; };
    float *gA = slot(G), *gB = slot(G + 1), *gzp = slot(G + 2), *scratch = slot(G + 3);
    float *d_nll = scratch + (int64_t)(W / 4) * S, *d_sdz = d_nll + n;
    double *d_stats = dscratch, *d_sums = dscratch + 128, *d_sg = dscratch + 136, *d_mix = dscratch + 256;
    const double cnt = (double)n * NF_PIXELS;
    const int sms = num_ctas_for(m);
    std::vector<int64_t> goff(m->layers.size() + 1);
    nf_grad_layout(m, goff.data());
    memset(grads_host, 0, sizeof(double) * (size_t)goff.back());
    std::vector<float> bn_all((size_t)G * 4 * W, 0.f);
    const int n_gd = nf_train_wide_grad_doubles(W), n_pf = nf_train_wide_param_floats(W);
    double* d_cg = nullptr;
    float* d_par = nullptr;
    NF_CUDA(cudaMallocAsync((void**)&d_cg, (size_t)n_gd * sizeof(double), stream));
    NF_CUDA(cudaMallocAsync((void**)&d_par, (size_t)n_pf * sizeof(float), stream));
    struct Free { double* a; float* b; cudaStream_t s; ~Free() { cudaFreeAsync(a, s); cudaFreeAsync(b, s); } } guard{d_cg, d_par, stream};
    // ------------------------------------------------------------ forward, keeping every op's input
    for (int g = 0; g < G; ++g) {
        const Group& gr = groups[g];
        const float* in = g == 0 ? x : slot(g - 1);
        float* bn = &bn_all[(size_t)g * 4 * W];
        if (!y && range_has_sdn(m, gr.lo, gr.hi)) return fail(NF_ERR_INVALID, "clean patch y is required by an sdn layer");
        if (gr.cl >= 0) {
            const Layer& L = m->layers[gr.cl];
            const WideRawView r(L.wraw.data(), W);
            if (batch_stats) {
                const float ident = 1.0f - L.raw.bn_eps;
                for (int k = 0; k < W; ++k) { bn[k] = 0.f; bn[2 * W + k] = 0.f; bn[W + k] = ident; bn[3 * W + k] = ident; }
                for (int stage = 1; stage <= 2; ++stage) {
                    NF_CUDA(cudaMemsetAsync(d_stats, 0, 2 * (size_t)W * sizeof(double), stream));
                    NfChainArgs a = {};
                    a.in = in; a.y = y; a.rows = rows; a.n = n; a.default_row = default_row; a.temp = 1.f;
                    a.bn_stats = d_stats; a.bn_stage = stage;
                    rc = launch_custom(m, gr.lo, gr.hi, true, a, gr.cl, bn, stream, true);
                    if (rc) return rc;
                    std::vector<double> h(2 * (size_t)W);
                    NF_CUDA(cudaMemcpyAsync(h.data(), d_stats, 2 * (size_t)W * sizeof(double), cudaMemcpyDeviceToHost, stream));
                    NF_CUDA(cudaStreamSynchronize(stream));
                    for (int k = 0; k < W; ++k) {
                        const double mean = h[k] / cnt;
                        double var = h[W + k] / cnt - mean * mean;
                        if (var < 0.0) var = 0.0;
                        bn[(stage - 1) * 2 * W + k] = (float)mean;
                        bn[(stage - 1) * 2 * W + W + k] = (float)var;
                    }
                }
            } else {
                memcpy(bn, r.bn1_mean, W * sizeof(float)); memcpy(bn + W, r.bn1_var, W * sizeof(float));
                memcpy(bn + 2 * W, r.bn2_mean, W * sizeof(float)); memcpy(bn + 3 * W, r.bn2_var, W * sizeof(float));
            }
            if (batch_stats_host) {
                int idx = 0;
                for (int l = 0; l < gr.cl; ++l) idx += m->layers[l].kind == L_COUPLING;
                memcpy(batch_stats_host + (size_t)4 * W * idx, bn, (size_t)4 * W * sizeof(float));
            }
        }
        NfChainArgs a = {};
        a.in = in; a.y = y; a.rows = rows; a.out = slot(g); a.n = n; a.default_row = default_row; a.temp = 1.f;
        a.logdet_in = g > 0 ? d_nll : nullptr;
        if (g + 1 == G) { a.nll = d_nll; a.sdz = d_sdz; } else a.logdet = d_nll;
        rc = launch_custom(m, gr.lo, gr.hi, true, a, gr.cl, gr.cl >= 0 ? bn : nullptr, stream, true);
        if (rc) return rc;
    }
    if (sums_host) {
        cudaError_t e = nf::launch_reduce(d_nll, d_sdz, n, d_sums, stream);
        if (e != cudaSuccess) return fail(NF_ERR_CUDA, "reduce launch: %s", cudaGetErrorString(e));
        NF_CUDA(cudaMemcpyAsync(sums_host, d_sums, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    // ------------------------------------------------------------ backward
    {
        cudaError_t e = nf::launch_train_prior(slot(G - 1), gA, n, sms, stream);
        if (e != cudaSuccess) return fail(NF_ERR_CUDA, "prior launch: %s", cudaGetErrorString(e));
    }
    std::vector<float> par((size_t)n_pf);
    std::vector<double> hg((size_t)n_gd);
    for (int g = G - 1; g >= 0; --g) {
        const Group& gr = groups[g];
        const float* zin = g == 0 ? x : slot(g - 1);
        const float* zout = slot(g);
        cudaError_t e = cudaSuccess;
        if (gr.cl >= 0) {
            const Layer& L = m->layers[gr.cl];
            const WideRawView r(L.wraw.data(), W);
            const float* bn = &bn_all[(size_t)g * 4 * W];
            const bool fused = gr.hi - gr.lo == 2;
            std::fill(par.begin(), par.end(), 0.f);
            for (int i = 0; i < 4; ++i)
                for (int o = 0; o < 4; ++o) par[i * 4 + o] = fused ? m->layers[gr.lo].a[o][i] : (i == o ? 1.f : 0.f);   // a[o][i] = A[i][o]
            par[16] = fused ? 1.f : 0.f;
            par[17] = L.raw.rescaling_scale;
            memcpy(&par[20], r.last_b, 16); memcpy(&par[24], r.last_logs, 16);
            float* q = &par[32];
            memcpy(q, r.l1_b, W * 4); q += W;
            memcpy(q, bn, W * 4); q += W;
            for (int k = 0; k < W; ++k) q[k] = (float)(1.0 / sqrt((double)bn[W + k] + (double)L.raw.bn_eps));
            q += W;
            memcpy(q, r.l2_b, W * 4); q += W;
            memcpy(q, bn + 2 * W, W * 4); q += W;
            for (int k = 0; k < W; ++k) q[k] = (float)(1.0 / sqrt((double)bn[3 * W + k] + (double)L.raw.bn_eps));
            q += W;
            memcpy(q, r.l1_w, (size_t)18 * W * 4); q += 18 * W;
            memcpy(q, r.l2_w, (size_t)W * W * 4); q += W * W;
            memcpy(q, r.last_w, (size_t)36 * (W + 1) * 4);
            NF_CUDA(cudaMemcpyAsync(d_par, par.data(), (size_t)n_pf * sizeof(float), cudaMemcpyHostToDevice, stream));
            NF_CUDA(cudaMemsetAsync(d_cg, 0, (size_t)n_gd * sizeof(double), stream));
            e = nf::launch_train_wide(W, 1, d_par, zin, gA, gzp, scratch, nullptr, n, nullptr, d_cg, sms, stream);
            if (e != cudaSuccess) return fail(NF_ERR_CUDA, "wide B1 launch: %s", cudaGetErrorString(e));
            const int oBN2 = n_gd - 4 * W, oBN1 = n_gd - 2 * W;
            NfBnTermsWide t2 = {}, t1 = {};
            std::vector<double> h(2 * (size_t)W);
            if (batch_stats) {
                NF_CUDA(cudaMemcpyAsync(h.data(), d_cg + oBN2, 2 * (size_t)W * sizeof(double), cudaMemcpyDeviceToHost, stream));
                NF_CUDA(cudaStreamSynchronize(stream));
                for (int k = 0; k < 2 * W; ++k) t2.v[k] = (float)(h[k] / cnt);
            }
            e = nf::launch_train_wide(W, 2, d_par, zin, nullptr, nullptr, scratch, nullptr, n, &t2, d_cg, sms, stream);
            if (e != cudaSuccess) return fail(NF_ERR_CUDA, "wide B2 launch: %s", cudaGetErrorString(e));
            if (batch_stats) {
                NF_CUDA(cudaMemcpyAsync(h.data(), d_cg + oBN1, 2 * (size_t)W * sizeof(double), cudaMemcpyDeviceToHost, stream));
                NF_CUDA(cudaStreamSynchronize(stream));
                for (int k = 0; k < 2 * W; ++k) t1.v[k] = (float)(h[k] / cnt);
            }
            e = nf::launch_train_wide(W, 3, d_par, zin, nullptr, gzp, scratch, gB, n, &t1, d_cg, sms, stream);
            if (e != cudaSuccess) return fail(NF_ERR_CUDA, "wide B3 launch: %s", cudaGetErrorString(e));
            NF_CUDA(cudaMemcpyAsync(hg.data(), d_cg, (size_t)n_gd * sizeof(double), cudaMemcpyDeviceToHost, stream));
            NF_CUDA(cudaStreamSynchronize(stream));
            // device block: A 16 | W1 | b1 | W2 | b2 | W3 | b3 | logs | scale ...  ->  host block: W1 .. scale (contiguous)
            memcpy(grads_host + goff[gr.cl], hg.data() + 16, (size_t)nf_train_wide_host_coupling(W) * sizeof(double));
            if (fused && m->layers[gr.lo].kind == L_CONV1X1) memcpy(grads_host + goff[gr.lo], hg.data(), 16 * sizeof(double));
        } else {
            const Layer& L = m->layers[gr.lo];
            if (L.kind == L_SCALE) {
                NfTrainScale T = {};
                for (int r = 0; r < NF_MAX_ROWS; ++r) { T.t[r][0] = L.table[r][0]; T.t[r][1] = L.table[r][1]; }
                T.is_sdn = L.scale_kind == NF_SCALE_SDN; T.full_sum = L.full_sum;
                NF_CUDA(cudaMemsetAsync(d_sg, 0, 2 * NF_MAX_ROWS * sizeof(double), stream));
                e = nf::launch_train_scale(zout, y, gA, gB, rows, default_row, n, T, d_sg, sms, stream);
                if (e != cudaSuccess) return fail(NF_ERR_CUDA, "scale backward launch: %s", cudaGetErrorString(e));
                NF_CUDA(cudaMemcpyAsync(grads_host + goff[gr.lo], d_sg, 2 * (size_t)L.n_rows * sizeof(double), cudaMemcpyDeviceToHost, stream));
                NF_CUDA(cudaStreamSynchronize(stream));
            } else {
                NfTrainMix M;
                for (int i = 0; i < 4; ++i)
                    for (int o = 0; o < 4; ++o) M.A[i][o] = L.a[o][i];
                NF_CUDA(cudaMemsetAsync(d_mix, 0, 16 * sizeof(double), stream));
                e = nf::launch_train_mix(zin, gA, gB, n, M, d_mix, sms, stream);
                if (e != cudaSuccess) return fail(NF_ERR_CUDA, "mix backward launch: %s", cudaGetErrorString(e));
                if (L.kind == L_CONV1X1) {
                    NF_CUDA(cudaMemcpyAsync(grads_host + goff[gr.lo], d_mix, 16 * sizeof(double), cudaMemcpyDeviceToHost, stream));
                    NF_CUDA(cudaStreamSynchronize(stream));
                }
            }
        }
        float* t = gA; gA = gB; gB = t;
    }
    NF_CUDA(cudaStreamSynchronize(stream));
    return NF_OK;
}

int nf_loss_and_grad(const nf_model* m, const float* x, const float* y, const int32_t* rows, int32_t default_row, int64_t n,
                     int batch_stats, float* workspace, double* dscratch, double* grads_host, float* batch_stats_host,
                     double* sums_host, void* stream_) {
    int rc = check_ready(m);
    if (rc) return rc;
    if (m->has_cond) return fail(NF_ERR_UNSUPPORTED, "the train-step kernels do not cover clean-image-conditioned couplings (legacy revnet2d models)");
    if (m->width != 4) {
        if (n <= 0 || !x || !workspace || !dscratch || !grads_host) return fail(NF_ERR_INVALID, "x, workspace, dscratch, grads_host are required and n > 0");
        if (default_row < 0 || default_row >= NF_MAX_ROWS) return fail(NF_ERR_INVALID, "default_row out of range");
        return loss_and_grad_wide(m, x, y, rows, default_row, n, batch_stats, workspace, dscratch, grads_host, batch_stats_host, sums_host,
                                  (cudaStream_t)stream_);
    }
    if (n <= 0 || !x || !workspace || !dscratch || !grads_host) return fail(NF_ERR_INVALID, "x, workspace, dscratch, grads_host are required and n > 0");
    if (default_row < 0 || default_row >= NF_MAX_ROWS) return fail(NF_ERR_INVALID, "default_row out of range");
    cudaStream_t stream = (cudaStream_t)stream_;
    const std::vector<Group> groups = make_groups(m);
    const int G = (int)groups.size();
    const int64_t S = n * NF_DIMS;
    auto slot = [&](int k) { return workspace + (int64_t)k * S; };
    float *gA = slot(G), *gB = slot(G + 1), *scratch = slot(G + 2), *gzp = slot(G + 3);
    float *d_nll = slot(G + 4), *d_sdz = d_nll + n;
    double *d_cg = dscratch, *d_stats = dscratch + 320, *d_sums = dscratch + 328, *d_sg = dscratch + 336;
    const double cnt = (double)n * NF_PIXELS;
    const int sms = num_ctas_for(m);
    std::vector<int64_t> goff(m->layers.size() + 1);
    nf_grad_layout(m, goff.data());
    memset(grads_host, 0, sizeof(double) * (size_t)goff.back());
    std::vector<float> bn_all((size_t)G * 16, 0.f);

    // ------------------------------------------------------------ forward, keeping every op's input
    for (int g = 0; g < G; ++g) {
        const Group& gr = groups[g];
        const float* in = g == 0 ? x : slot(g - 1);
        float* out = slot(g);
        float* bn = &bn_all[(size_t)g * 16];
        if (!y && range_has_sdn(m, gr.lo, gr.hi)) return fail(NF_ERR_INVALID, "clean patch y is required by an sdn layer");
        if (gr.cl >= 0) {
            const Layer& L = m->layers[gr.cl];
            if (batch_stats) {
                for (int k = 0; k < 4; ++k) { bn[k] = 0.f; bn[8 + k] = 0.f; bn[4 + k] = bn[12 + k] = 1.0f - L.raw.bn_eps; }
                for (int stage = 1; stage <= 2; ++stage) {
                    NfModelParams mp;
                    float ldjc = 0.f;
                    { std::lock_guard<std::mutex> lock(m->prog_mu); rc = build_program(m, gr.lo, gr.hi, &mp, &ldjc, gr.cl, bn); }
                    if (rc) return rc;
                    NF_CUDA(cudaMemsetAsync(d_stats, 0, 8 * sizeof(double), stream));
                    NfChainArgs a = {};
                    a.in = in; a.y = y; a.rows = rows; a.n = n; a.default_row = default_row; a.temp = 1.f;
                    a.first_layer = 0; a.last_layer = mp.n_layers; a.bn_stats = d_stats; a.bn_stage = stage;
                    cudaError_t e = nf::launch_chain(mp, a, true, sms, m->warps_per_cta, stream);
                    if (e != cudaSuccess) return fail(NF_ERR_CUDA, "probe launch: %s", cudaGetErrorString(e));
                    double h[8];
                    NF_CUDA(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, stream));
                    NF_CUDA(cudaStreamSynchronize(stream));
                    for (int k = 0; k < 4; ++k) {
                        const double mean = h[k] / cnt;
                        double var = h[4 + k] / cnt - mean * mean;
                        if (var < 0.0) var = 0.0;
                        bn[(stage - 1) * 8 + k] = (float)mean;
                        bn[(stage - 1) * 8 + 4 + k] = (float)var;
                    }
                }
            } else {
                memcpy(bn, L.raw.bn1_mean, 16); memcpy(bn + 4, L.raw.bn1_var, 16);
                memcpy(bn + 8, L.raw.bn2_mean, 16); memcpy(bn + 12, L.raw.bn2_var, 16);
            }
            if (batch_stats_host) {
                int idx = 0;
                for (int l = 0; l < gr.cl; ++l) idx += m->layers[l].kind == L_COUPLING;
                memcpy(batch_stats_host + 16 * idx, bn, 16 * sizeof(float));
            }
        }
        NfModelParams mp;
        float ldjc = 0.f;
        { std::lock_guard<std::mutex> lock(m->prog_mu); rc = build_program(m, gr.lo, gr.hi, &mp, &ldjc, gr.cl, gr.cl >= 0 ? bn : nullptr); }
        if (rc) return rc;
        NfChainArgs a = {};
        a.in = in; a.y = y; a.rows = rows; a.out = out; a.n = n; a.default_row = default_row; a.temp = 1.f;
        a.first_layer = 0; a.last_layer = mp.n_layers; a.ldj_const = ldjc;
        a.logdet_in = g > 0 ? d_nll : nullptr;
        if (g + 1 == G) { a.nll = d_nll; a.sdz = d_sdz; } else a.logdet = d_nll;
        cudaError_t e = nf::launch_chain(mp, a, true, sms, m->warps_per_cta, stream);
        if (e != cudaSuccess) return fail(NF_ERR_CUDA, "forward launch: %s", cudaGetErrorString(e));
    }
    if (sums_host) {
        cudaError_t e = nf::launch_reduce(d_nll, d_sdz, n, d_sums, stream);
        if (e != cudaSuccess) return fail(NF_ERR_CUDA, "reduce launch: %s", cudaGetErrorString(e));
        NF_CUDA(cudaMemcpyAsync(sums_host, d_sums, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    // ------------------------------------------------------------ backward
    {
        cudaError_t e = nf::launch_train_prior(slot(G - 1), gA, n, sms, stream);
        if (e != cudaSuccess) return fail(NF_ERR_CUDA, "prior launch: %s", cudaGetErrorString(e));
    }
    for (int g = G - 1; g >= 0; --g) {
        const Group& gr = groups[g];
        const float* zin = g == 0 ? x : slot(g - 1);
        const float* zout = slot(g);
        cudaError_t e = cudaSuccess;
        if (gr.cl >= 0) {
            NfTrainCoupling P;
            fill_train_coupling(m, gr, &bn_all[(size_t)g * 16], &P);
            NF_CUDA(cudaMemsetAsync(d_cg, 0, NF_G_COUPLING_DOUBLES * sizeof(double), stream));
            e = nf::launch_train_b1(P, zin, gA, gzp, scratch, n, d_cg, sms, stream);
            if (e != cudaSuccess) return fail(NF_ERR_CUDA, "B1 launch: %s", cudaGetErrorString(e));
            double h[8];
            NfBnTerms t2 = {}, t1 = {};
            if (batch_stats) {
                NF_CUDA(cudaMemcpyAsync(h, d_cg + NF_G_BN2, sizeof(h), cudaMemcpyDeviceToHost, stream));
                NF_CUDA(cudaStreamSynchronize(stream));
                for (int k = 0; k < 8; ++k) t2.v[k] = (float)(h[k] / cnt);
            }
            e = nf::launch_train_b2(P, zin, scratch, n, t2, d_cg, sms, stream);
            if (e != cudaSuccess) return fail(NF_ERR_CUDA, "B2 launch: %s", cudaGetErrorString(e));
            if (batch_stats) {
                NF_CUDA(cudaMemcpyAsync(h, d_cg + NF_G_BN1, sizeof(h), cudaMemcpyDeviceToHost, stream));
                NF_CUDA(cudaStreamSynchronize(stream));
                for (int k = 0; k < 8; ++k) t1.v[k] = (float)(h[k] / cnt);
            }
            e = nf::launch_train_b3(P, zin, scratch, gzp, gB, n, t1, d_cg, sms, stream);
            if (e != cudaSuccess) return fail(NF_ERR_CUDA, "B3 launch: %s", cudaGetErrorString(e));
            double hg[NF_G_COUPLING_DOUBLES];
            NF_CUDA(cudaMemcpyAsync(hg, d_cg, sizeof(hg), cudaMemcpyDeviceToHost, stream));
            NF_CUDA(cudaStreamSynchronize(stream));
            memcpy(grads_host + goff[gr.cl], hg + NF_G_W1, NF_G_HOST_COUPLING * sizeof(double));   // W1..scale are contiguous
            if (gr.hi - gr.lo == 2 && m->layers[gr.lo].kind == L_CONV1X1) memcpy(grads_host + goff[gr.lo], hg + NF_G_A, 16 * sizeof(double));
        } else {
            const Layer& L = m->layers[gr.lo];
            if (L.kind == L_SCALE) {
                NfTrainScale T = {};
                for (int r = 0; r < NF_MAX_ROWS; ++r) { T.t[r][0] = L.table[r][0]; T.t[r][1] = L.table[r][1]; }
                T.is_sdn = L.scale_kind == NF_SCALE_SDN; T.full_sum = L.full_sum;
                NF_CUDA(cudaMemsetAsync(d_sg, 0, 2 * NF_MAX_ROWS * sizeof(double), stream));
                e = nf::launch_train_scale(zout, y, gA, gB, rows, default_row, n, T, d_sg, sms, stream);
                if (e != cudaSuccess) return fail(NF_ERR_CUDA, "scale backward launch: %s", cudaGetErrorString(e));
                NF_CUDA(cudaMemcpyAsync(grads_host + goff[gr.lo], d_sg, 2 * (size_t)L.n_rows * sizeof(double), cudaMemcpyDeviceToHost, stream));
                NF_CUDA(cudaStreamSynchronize(stream));
            } else {   // stand-alone conv1x1 / permutation
                NfTrainMix M;
                for (int i = 0; i < 4; ++i)
                    for (int o = 0; o < 4; ++o) M.A[i][o] = L.a[o][i];
                NF_CUDA(cudaMemsetAsync(d_cg, 0, 16 * sizeof(double), stream));
                e = nf::launch_train_mix(zin, gA, gB, n, M, d_cg, sms, stream);
                if (e != cudaSuccess) return fail(NF_ERR_CUDA, "mix backward launch: %s", cudaGetErrorString(e));
                if (L.kind == L_CONV1X1) {
                    NF_CUDA(cudaMemcpyAsync(grads_host + goff[gr.lo], d_cg, 16 * sizeof(double), cudaMemcpyDeviceToHost, stream));
                    NF_CUDA(cudaStreamSynchronize(stream));
                }
            }
        }
        float* t = gA; gA = gB; gB = t;
    }
    NF_CUDA(cudaStreamSynchronize(stream));
    return NF_OK;
}

int nf_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(NF_ERR_INVALID, "ptr is null");
    NF_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return NF_OK;
}
int nf_host_free(void* ptr) {
    if (ptr) NF_CUDA(cudaFreeHost(ptr));
    return NF_OK;
}

}  // extern "C"
