// Parameter blocks and gradient layout of the backward kernels (nf_train.cu); host side in nf_api.cu.
#pragma once
#include <cuda_runtime.h>
#include "nf_params.h"

// Raw (un-folded) parameters of one coupling as the reference stores them, plus the BatchNorm statistics in
// force for this step (batch statistics when is_training, moving statistics otherwise) as mean / 1/sqrt(var+eps).
struct NfTrainCoupling {
    float A[4][4];            // fused 1x1 conv, [in][out] (identity when has_mix == 0)
    float w1[3][3][2][4];     // TF layout [kh][kw][in][out]
    float w2[4][4];           // [in][out]
    float w3[3][3][5][4];     // in = 4 hidden channels + edge indicator
    float b1[4], m1[4], is1[4];
    float b2[4], m2[4], is2[4];
    float b3[4], logs[4];
    float scale;
    int32_t has_mix;
    float pad_[2];
};

struct NfBnTerms { float v[8]; };    // [0..3] = mean(g_hat), [4..7] = mean(g_hat * x_hat); zeros for moving statistics

struct NfTrainScale {
    float t[NF_MAX_ROWS][2];          // sdn: (a, b); gain: (g, -)
    int32_t is_sdn, full_sum;
};

struct NfTrainMix { float A[4][4]; };   // [in][out]

// device gradient block of one coupling (doubles)
#define NF_G_A 0        // 16  d loss / d A[in][out]
#define NF_G_W1 16      // 72
#define NF_G_B1 88      // 4
#define NF_G_W2 92      // 16
#define NF_G_B2 108     // 4
#define NF_G_W3 112     // 180
#define NF_G_B3 292     // 4
#define NF_G_LOGS 296   // 4
#define NF_G_SCALE 300  // 1
#define NF_G_BN2 304    // 8: sum g_c2hat, sum g_c2hat * c2hat
#define NF_G_BN1 312    // 8
#define NF_G_COUPLING_DOUBLES 320
// host gradient block of one coupling: [W1 72][b1 4][W2 16][b2 4][W3 180][b3 4][logs 4][scale 1]
#define NF_G_HOST_COUPLING 285

// ---- coupling nets wider than 4 (nf_train_wide.cu, widths 8 / 16 / 32): one CTA per patch
// device parameter block of one coupling (floats), raw TF layouts:
//   A [4][4] ([in][out]) | META has_mix, rescaling_scale, 0, 0 | b3 [4] | logs [4] | pad to 32
//   b1 [W] | m1 [W] | is1 [W] | b2 [W] | m2 [W] | is2 [W] | w1 [9][2][W] | w2 [W][W] ([in][out]) | w3 [9][W+1][4]
// device gradient block (doubles):
//   A 16 | W1 18W | b1 W | W2 W*W | b2 W | W3 36(W+1) | b3 4 | logs 4 | scale 1 (+3 pad) | BN2 sums 2W | BN1 sums 2W
// host gradient block of one coupling: [W1][b1][W2][b2][W3][b3][logs][scale] contiguous from W1, as for width 4
constexpr int nf_train_wide_param_floats(int W) { return 32 + 6 * W + 18 * W + W * W + 36 * (W + 1); }
constexpr int nf_train_wide_grad_doubles(int W) { return 16 + 18 * W + W + W * W + W + 36 * (W + 1) + 4 + 4 + 4 + 4 * W; }
constexpr int nf_train_wide_host_coupling(int W) { return 18 * W + W + W * W + W + 36 * (W + 1) + 4 + 4 + 1; }
struct NfBnTermsWide { float v[64]; };   // [0..W) = mean(g_hat), [W..2W) = mean(g_hat * x_hat); zeros for moving statistics

namespace nf {
bool train_wide_width_supported(int W);
// pass 1 / 2 / 3 = B1 / B2 / B3 (see nf_train_wide.cu); scratch: n * 1024 * W floats
cudaError_t launch_train_wide(int W, int pass, const float* params, const float* zin, const float* gout, float* gzp, float* scratch, float* gin,
                              long long n, const NfBnTermsWide* bn, double* grads, int num_sms, cudaStream_t s);
cudaError_t launch_train_b1(const NfTrainCoupling& P, const float* zin, const float* gout, float* gzp, float* scratch, long long n,
                            double* grads, int num_sms, cudaStream_t s);
cudaError_t launch_train_b2(const NfTrainCoupling& P, const float* zin, float* scratch, long long n, const NfBnTerms& bn2,
                            double* grads, int num_sms, cudaStream_t s);
cudaError_t launch_train_b3(const NfTrainCoupling& P, const float* zin, const float* scratch, const float* gzp, float* gin,
                            long long n, const NfBnTerms& bn1, double* grads, int num_sms, cudaStream_t s);
cudaError_t launch_train_scale(const float* zout, const float* y, const float* gout, float* gin, const int* rows, int default_row,
                               long long n, const NfTrainScale& T, double* grads, int num_sms, cudaStream_t s);
cudaError_t launch_train_mix(const float* zin, const float* gout, float* gin, long long n, const NfTrainMix& M, double* grads,
                             int num_sms, cudaStream_t s);
cudaError_t launch_train_prior(const float* z, float* g, long long n, int num_sms, cudaStream_t s);

// launch descriptor of one coupling: offsets into the flat variable array (constant from step to step)
struct TdCoupling {
    int32_t off_w1, off_b1, off_w2, off_b2, off_w3, off_b3, off_logs, off_scale;
    int32_t off_bn[4];          // moving mean1, var1, mean2, var2
    int32_t has_mix;            // A comes from the derived-matrix buffer
    int32_t batch_stats;
    float bn_eps;
    int32_t pad_;
};

// ---- small-batch chain with batch-statistics BatchNorm as one cooperative kernel (nf_trainer.cu: td_bs_chain_kernel)
struct BsOp {
    int32_t kind;               // 0 = coupling (+ the mix in front of it), 1 = scale layer
    int32_t cidx, sidx;         // statistics / matrix index, table index
    int32_t is_sdn, full_sum;
    int32_t pad_[3];
    TdCoupling d;               // offsets into the flat parameter array
};
struct BsArgs {
    const BsOp* ops;
    int32_t n_ops, direction;
    const float *vars, *Amat, *tables;          // Amat: [n_couplings][in][out], A (direction 0) or A^-1 (direction 1)
    double* stats;                              // [n_couplings][16], zeroed
    const double* consts;                       // [0] = constant log-det of the chain
    const float *in, *y;
    const int32_t* rows;
    int32_t default_row, pad_;
    long long n;
    float *out, *ld, *nll, *sdz, *logdet;       // ld: [n] zeroed scratch (may alias nll or logdet); the rest as NfChainArgs
    float temp, pad2_;
    unsigned long long seed, offset, patch_base;
};

int bs_small_capacity(int sm_count);                       // patches that can own a co-resident CTA (0: unavailable)
cudaError_t launch_bs_small(const BsArgs& a, int capacity, cudaStream_t s);   // grid = min(n, capacity); beyond: grid-stride
}  // namespace nf
