// Wide coupling nets (width 8 / 16 / 32): parameter blob layout and launcher shared by nf_api.cu (host folding)
// and nf_wide.cu (device code).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_params.h"

// One folded coupling inside the model's device blob (floats; every offset is a multiple of 4 -> float4 loads).
// CIN = input channels of the coupling net, COUT = its output channels:
//   AffineCoupling           (layers.py:251-375)                  net(x0)            CIN 2, COUT 4   mode 0
//   AffineCouplingCondXY[G]  (noise_flow_layers/AffineCouplingCondXY.py:45-79)   net(concat(x0, yy))  CIN 6, COUT 4   mode 1
//   AffineCouplingCondY[G]   (noise_flow_layers/AffineCouplingCondY.py:44-72)    net(yy), all four channels transformed
//                                                                                       CIN 4, COUT 8   mode 2
//   A     [4][4]      mix, inverse direction, [o][i]            (identity when has_mix == 0)
//   AINV  [4][4]      mix, forward direction, [o][i]
//   META  has_mix, rescaling_scale, mode, 0
//   B3    [3][3][COUT] (b3 + edge-indicator taps) * exp(3 logs) by (row class, column class)   (72 floats reserved)
//   B1    [W]         (b1 - mean1) / sqrt(var1 + eps)
//   B2    [W]
//   W1    [9][W][CIN] conv 3x3 CIN -> W, BN-1 folded, [tap][o][i]
//   W2    [W][W]      conv 1x1 W -> W, BN-2 folded, [o][i]
//   W3    [9][W][COUT] conv 3x3 W -> COUT (zero-padded h2), * exp(3 logs), [tap][i][o]
struct NfWideLayoutG {
    int W, CIN, COUT;
    __host__ __device__ constexpr NfWideLayoutG(int w, int cin, int cout) : W(w), CIN(cin), COUT(cout) {}
    static constexpr int A = 0, AINV = 16, META = 32, B3 = 36, B1 = 108;
    __host__ __device__ constexpr int b2() const { return B1 + W; }
    __host__ __device__ constexpr int w1() const { return b2() + W + ((4 - (2 * W) % 4) % 4); }
    __host__ __device__ constexpr int w2() const { return w1() + 9 * W * CIN + ((4 - (9 * W * CIN) % 4) % 4); }
    __host__ __device__ constexpr int w3() const { return w2() + W * W; }
    __host__ __device__ constexpr int size() const { return w3() + 9 * W * COUT; }
};
template <int W>
struct NfWideLayout {      // the plain AffineCoupling (CIN 2, COUT 4)
    static constexpr NfWideLayoutG G = NfWideLayoutG(W, 2, 4);
    static constexpr int A = 0, AINV = 16, META = 32, B3 = 36, B1 = NfWideLayoutG::B1, B2 = G.b2(), W1 = G.w1(), W2 = G.w2(),
                         W3 = G.w3(), SIZE = G.size();
};
inline int nf_wide_coupling_floats(int W, int cin = 2, int cout = 4) { return NfWideLayoutG(W, cin, cout).size(); }
#define NF_COUPLING_MODE_X 0
#define NF_COUPLING_MODE_XY 1
#define NF_COUPLING_MODE_Y 2
#define NF_WIDE_SCALE_FLOATS (NF_MAX_ROWS * 4)   // scale op: NfScaleP::t
#define NF_WIDE_MIX_FLOATS 32                    // stand-alone mix op: a[4][4], ainv[4][4], [o][i]

#define NF_WIDE_FLAG_COND 1      // the program has clean-image-conditioned couplings: stage the clean patch in shared memory
struct NfWideProgram {           // by-value kernel argument; the parameters themselves live in the device blob
    int32_t width;
    int32_t n_layers;
    int32_t flags;
    int32_t blob_floats;         // size of the blob (tensor-core kernels: a multiple of NF_TMA_ROW_FLOATS, coupling blocks row-aligned)
    int32_t op[NF_MAX_LAYERS];   // NfKernelOp, data -> latent order
    int32_t off[NF_MAX_LAYERS];  // float offset of the op's block in the blob
};

// ---- tensor-core kernel (nf_wide_tc.cu): one coupling = one block, laid out exactly as it sits in shared memory so that
// the TMA engine copies it verbatim (byte offsets, every one a multiple of 128):
//   header  128 floats : A [4][4] | AINV [4][4] | META (has_mix, rescaling_scale, 0, 0) | B3 [3][3][4] edge-bias table
//   B1   [64/8][W][8]   bf16   conv-1, K rows: W1_hi (18) | W1_hi (18) | W1_lo (18) | b1_hi | b1_lo | 0 x 8   (BN-1 folded)
//   B2   [W/8][2W][8]   bf16   conv-2, N rows 0..W-1: W2_hi[o][k], rows W..2W-1: W2_lo[o][k]                  (BN-2 folded)
//   BB2  [16/8][W][8]   bf16   conv-2 bias: K row 6 = b2_hi, row 7 = b2_lo (the constant-one slots of A1's K chunk 3)
//   B3   [W/8][96][8]   bf16   conv-3 as a 1x1 GEMM, N row dy*16 + dx*4 + o: W3_hi ; row 48 + ...: W3_lo      (* exp(3 logs))
// "K-major, no swizzle": element (n, k) of an [K/8][N][8] matrix at ((k / 8) * N + n) * 8 + k % 8.
// Widths 256 / 512 (nf_wide_tcs.cu) use the same offsets with the matrices cut into the blocks its weight ring carries:
//   B1 [W/64][8][64][8], B2 [W/256 passes][W/64][8][512][8] (rows 0..255 hi, 256..511 lo of the pass's outputs),
//   BB2 [W/256][2][256][8], B3 unchanged.
struct NfWideTcLayout {
    static constexpr int H_A = 0, H_AINV = 16, H_META = 32, H_B3 = 36, HDR_FLOATS = 128;
    __host__ __device__ static constexpr int off_b1() { return HDR_FLOATS * 4; }
    __host__ __device__ static constexpr int off_b2(int W) { return off_b1() + 128 * W; }
    __host__ __device__ static constexpr int off_bb2(int W) { return off_b2(W) + 4 * W * W; }
    __host__ __device__ static constexpr int off_b3(int W) { return off_bb2(W) + 32 * W; }
    __host__ __device__ static constexpr int block_bytes(int W) { return off_b3(W) + 192 * W; }
};
inline int nf_wide_tc_coupling_floats(int W) { return NfWideTcLayout::block_bytes(W) / 4; }

// The tensor-core kernels fetch their weight blocks with TMA tensor copies (cp.async.bulk.tensor.2d): the blob is described
// to the TMA engine as a 2-D tensor of rows of 64 floats (256 B), one box = NF_TMA_ROW_FLOATS x box_rows.
#define NF_TMA_ROW_FLOATS 64
namespace nf {
inline int wide_tc_box_rows(int width) { return width == 32 ? 62 : (width == 64 ? 77 : (width == 128 ? 217 : 32)); }
bool wide_width_supported(int width);
cudaError_t launch_wide_cond(const NfWideProgram& prog, const float* blob, const NfChainArgs& args, bool inverse, unsigned grid,
                             cudaStream_t stream);
bool wide_tc_width_supported(int width);
cudaError_t launch_chain_wide_tc(const NfWideProgram& prog, const float* blob, const NfChainArgs& args, bool inverse, int num_sms,
                                 cudaStream_t stream);
cudaError_t launch_chain_wide_tcs(const NfWideProgram& prog, const float* blob, const NfChainArgs& args, bool inverse, int num_sms,
                                  cudaStream_t stream);
cudaError_t launch_chain_wide(const NfWideProgram& prog, const float* blob, const NfChainArgs& args, bool inverse, int num_sms,
                              cudaStream_t stream);
}  // namespace nf
