// Kernel-ready ("folded") model parameters shared by the host packer (nf_api.cu) and the
// device code (nf_kernels.cu).  The whole struct is passed BY VALUE as a __grid_constant__
// kernel parameter (constant bank 0, broadcast through the uniform datapath: LDCU + FFMA2 with
// uniform-register operands), so a launch is self-contained and the C-ABI stays re-entrant:
// no __constant__ symbols, no per-handle device allocations.
#pragma once
#include <stdint.h>

#define NF_PATCH_H 32
#define NF_PATCH_W 32
#define NF_PATCH_C 4
#define NF_PIXELS (NF_PATCH_H * NF_PATCH_W)
#define NF_DIMS (NF_PIXELS * NF_PATCH_C)

#define NF_MAX_COUPLINGS 16   // fused (1x1-mix +) affine-coupling slots with compile-time parameter offsets
#define NF_MAX_MIX 4          // stand-alone 1x1 convs / permutations (not followed by a coupling)
#define NF_MAX_SCALE 4        // sdn*/gain* scale layers
#define NF_MAX_ROWS 32        // rows of the per-(camera, ISO) conditioning table
#define NF_MAX_LAYERS 48

enum NfKernelOp : int32_t {
    NF_KOP_COUPLING = 1,   // slot -> ModelParams::cp
    NF_KOP_MIX = 2,        // slot -> ModelParams::mix
    NF_KOP_SDN = 3,        // slot -> ModelParams::sc, scale^2 = a*y + b
    NF_KOP_GAIN = 4        // slot -> ModelParams::sc, scale = g
};

// One affine coupling (reference layers.py:251-375) with its real_nvp_conv_template
// (layers.py:452-498) folded for moving-statistics BatchNorm, optionally preceded (data->latent
// direction) by the invertible 1x1 conv / channel permutation that the reference places before
// every 'unc' coupling (noise_flow_model.py:79-104).  Weight layouts are [.. out][in] so that an
// (in0,in1) pair is one float2 -> packed fma.rn.f32x2 over input-channel pairs.
struct alignas(16) NfCouplingP {
    float a[4][4];          // mix, inverse direction:  out[o] = sum_i z[i] * a[o][i]   (= A[i][o])
    float ainv[4][4];       // mix, forward direction:  out[o] = sum_i z[i] * ainv[o][i] (= A_inv[i][o])
    float w1[3][3][4][2];   // conv 3x3 SAME 2->4, BN1 folded: [dy][dx][o][i]
    float w2[4][4];         // conv 1x1 4->4, BN2 folded: [o][i]
    float w3[3][3][4][4];   // conv 3x3 (edge-padded, VALID) 4->4, * exp(3*logs): [dy][dx][o][i]
    float b1[4];            // (b1 - mean1) / sqrt(var1 + eps)
    float b2[4];
    float b3[3][3][4];      // [row class][col class][o]: (b3 + edge-indicator taps) * exp(3*logs)
    float scale;            // rescaling_scale (layers.py:271-273)
    int32_t has_mix;
    float pad_[2];
};

struct NfMixP {             // stand-alone 4x4 channel mix
    float a[4][4];          // inverse direction, [o][i]
    float ainv[4][4];       // forward direction, [o][i]
};

// Scale layer table, one row per conditioning class (camera, ISO).
//   SDN : t[row] = {a, b, 0, 0}                scale = sqrt(a*y + b)
//   GAIN: t[row] = {g, 1/g, ldj_inverse, 0}    scale = g, ldj_inverse = -4096*log g (or -log g: quirk)
struct NfScaleP {
    float t[NF_MAX_ROWS][4];
};

struct NfModelParams {
    int32_t n_layers;
    int32_t n_rows;
    int32_t pad_[2];
    int32_t op[NF_MAX_LAYERS];     // NfKernelOp, in data->latent (inverse) order (int32: uniform LDCU indexing)
    int32_t slot[NF_MAX_LAYERS];
    NfCouplingP cp[NF_MAX_COUPLINGS];
    NfMixP mix[NF_MAX_MIX];
    NfScaleP sc[NF_MAX_SCALE];
};

// Per-launch arguments of the fused chain kernel.
struct NfChainArgs {
    const float* in;          // inverse: x (data)   forward: z (latent) or eps, may be null -> Philox
    const float* y;           // clean patch (conditioning); may be null if no SDN layer is run
    const int32_t* rows;      // per-patch conditioning row, or null -> default_row
    float* out;               // inverse: z (may be null)   forward: x
    float* logdet;            // [n] accumulated log-det (may be null)
    float* nll;               // [n] -(logdet + log N(z;0,I))  (inverse + prior; may be null)
    float* sdz;               // [n] sqrt(var(z))              (may be null)
    long long n;
    unsigned long long seed, offset;   // Philox key / stream offset (forward with in == null)
    unsigned long long patch_base;     // Philox patch counter of element 0 (sharded / chunked calls)
    int32_t first_layer, last_layer;   // layer range [first, last) in inverse order
    int32_t default_row;
    float ldj_const;          // host-computed constant log-det of the range (1x1 convs)
    float temp;               // forward: z = in * temp
    // batch-statistics BatchNorm support (layer-by-layer execution, see nf_chain_batch_stats in nf_api.cu)
    const float* logdet_in;   // [n] log-det accumulated by earlier launches (may be null)
    double* bn_stats;         // double[8]: per-channel sum and sum of squares of a pre-BN activation
    int32_t bn_stage;         // 0 = normal run; 1 / 2 = the LAST op of the range is a coupling: accumulate the
                              // statistics of its conv-1 / conv-2 output (before BatchNorm) and stop there
    int32_t pad_;
};
