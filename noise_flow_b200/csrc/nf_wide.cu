// Fused Noise Flow chain for coupling nets wider than the shipped width 4 (real_nvp_conv_template(width),
// layers.py:452-498; `--width` in sidd/ArgParser.py:43, "for Noise Flow it is 32": job_noise_flow.sh:19).
//
// ONE CTA OWNS ONE PATCH.  At width W the hidden activation of a coupling net is 34 x 34 x W floats (148 KB at
// W = 32) -- it no longer fits a warp's share of the SM, so the whole CTA (16 warps; warp w owns image rows
// w, w+16; lane = column) works on one patch that stays resident in shared memory across the whole chain:
//   z   [32][32] float4                        the patch
//   h2  [34][W/4][34] float4                   padded hidden image, channel groups of 4 as planes so that the 32
//                                              lanes of a warp read 32 consecutive float4 (conflict-free)
//   w   one coupling's folded parameters       staged from the model's device blob at every layer
// A coupling is two CTA-wide phases: P1 conv3x3(2->W)+BN+ReLU -> conv1x1(W->W)+BN+ReLU per pixel in registers,
// h2 to shared memory; P2 edge-padded conv3x3(W->4) -> tanh/exp affine update + log-det.  BatchNorm (moving
// statistics), exp(3 logs) and the edge-indicator channel are folded on the host exactly as for width 4
// (nf_api.cu fold_wide).  The binding roof is the FP32 pipe: 86 W + 18 W + W^2 + 36 W MAC per pixel and coupling.
#include <cuda_runtime.h>
#include <stdint.h>
#include "nf_kernels.h"
#include "nf_params.h"
#include "nf_rng.cuh"
#include "nf_wide.h"
#include "nf_wide_impl.cuh"

namespace nf {

template <int W, bool INV>
__global__ void __launch_bounds__(WIDE_THREADS, 1)
nf_wide_chain_kernel(const NfWideProgram prog, const float* __restrict__ blob, const NfChainArgs a) {
    wide_chain_body<W, false>(prog, blob, a, INV);
}

template <int W>
static cudaError_t launch_wide_w(const NfWideProgram& prog, const float* blob, const NfChainArgs& a, bool inverse, int num_sms,
                                 cudaStream_t stream) {
    const size_t smem = sizeof(WideSmem<W>);
    static bool attr_done_dev[NF_MAX_DEVICES] = {};   // per device (and per width: one instance of this template each)
    bool& attr_done = attr_done_dev[device_slot()];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(nf_wide_chain_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(nf_wide_chain_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const long long grid = a.n < (long long)num_sms ? a.n : (long long)num_sms;   // persistent: one CTA per SM
    if (prog.flags & NF_WIDE_FLAG_COND) return launch_wide_cond(prog, blob, a, inverse, (unsigned)grid, stream);
    if (inverse) nf_wide_chain_kernel<W, true><<<(unsigned)grid, WIDE_THREADS, smem, stream>>>(prog, blob, a);
    else nf_wide_chain_kernel<W, false><<<(unsigned)grid, WIDE_THREADS, smem, stream>>>(prog, blob, a);
    return cudaGetLastError();
}

bool wide_width_supported(int width) { return width == 4 || width == 8 || width == 16 || width == 32; }

cudaError_t launch_chain_wide(const NfWideProgram& prog, const float* blob, const NfChainArgs& a, bool inverse, int num_sms,
                              cudaStream_t stream) {
    switch (prog.width) {
        case 4: return launch_wide_w<4>(prog, blob, a, inverse, num_sms, stream);
        case 8: return launch_wide_w<8>(prog, blob, a, inverse, num_sms, stream);
        case 16: return launch_wide_w<16>(prog, blob, a, inverse, num_sms, stream);
        case 32: return launch_wide_w<32>(prog, blob, a, inverse, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace nf
