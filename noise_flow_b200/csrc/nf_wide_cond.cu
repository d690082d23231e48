// CTA-per-patch chain kernel for programs with clean-image-conditioned couplings -- the reference's legacy `revnet2d`
// models (noise_flow_model.py:237-392): AffineCouplingCondY / CondYG / CondXY / CondXYG
// (noise_flow_layers/AffineCouplingCond*.py), reached when hps.arch is unset.  Same device code as nf_wide.cu
// (nf_wide_impl.cuh) with the clean patch staged in shared memory and the direction chosen at run time; its own
// translation unit because of the 36 coupling instances (3 modes x 3 probe stages x 4 widths).
#include "nf_wide_impl.cuh"

namespace nf {

template <int W>
__global__ void __launch_bounds__(WIDE_THREADS, 1)
nf_wide_cond_kernel(const NfWideProgram prog, const float* __restrict__ blob, const NfChainArgs a, const int inverse) {
    wide_chain_body<W, true>(prog, blob, a, inverse != 0);
}

template <int W>
static cudaError_t launch_cond_w(const NfWideProgram& prog, const float* blob, const NfChainArgs& a, bool inverse, unsigned grid, cudaStream_t stream) {
    const size_t smem = sizeof(WideSmem<W>);
    cudaError_t e = cudaFuncSetAttribute(nf_wide_cond_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device: every launch
    if (e != cudaSuccess) return e;
    nf_wide_cond_kernel<W><<<grid, WIDE_THREADS, smem, stream>>>(prog, blob, a, inverse ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_wide_cond(const NfWideProgram& prog, const float* blob, const NfChainArgs& a, bool inverse, unsigned grid, cudaStream_t stream) {
    if (!a.y) return cudaErrorInvalidValue;      // the coupling nets read the clean patch
    switch (prog.width) {
        case 4: return launch_cond_w<4>(prog, blob, a, inverse, grid, stream);
        case 8: return launch_cond_w<8>(prog, blob, a, inverse, grid, stream);
        case 16: return launch_cond_w<16>(prog, blob, a, inverse, grid, stream);
        case 32: return launch_cond_w<32>(prog, blob, a, inverse, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace nf
