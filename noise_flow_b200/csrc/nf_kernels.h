// Internal launcher interface between nf_api.cu (C-ABI, host logic) and nf_kernels.cu (device code).
#pragma once
#include <cuda_runtime.h>
#include "nf_params.h"

#ifndef NF_Z_IN_TMEM
#define NF_Z_IN_TMEM 1   // resident patches live in tensor memory (16 per SM); 0: in shared memory (12 per SM)
#endif
#if NF_Z_IN_TMEM
#define NF_WARP_SMEM_BYTES (2 * 34 * 16 + 2 * 34 * 8)                    // row rings only: 1632 B per resident patch
#define NF_MAX_WARPS_PER_CTA 16                                          // 16 x 128 columns = the 512 TMEM columns
#else
#define NF_WARP_SMEM_BYTES (NF_PIXELS * 16 + 2 * 34 * 16 + 2 * 34 * 8)   // 18016 B per resident patch
#define NF_MAX_WARPS_PER_CTA 12
#endif
#define NF_MAX_CTA_THREADS (NF_MAX_WARPS_PER_CTA * 32)
#define NF_MAX_CTA_SMEM (NF_MAX_WARPS_PER_CTA * NF_WARP_SMEM_BYTES)

// tensor-core path (nf_tc.cu)
#define NF_TC_GROUPS 3   // patches in flight per CTA (4 warps each); 3 x 144 TMEM columns <= 512
#define NF_TC_SLOTS 8    // couplings whose B tiles are resident in shared memory

#define NF_MAX_DEVICES 64
namespace nf {
// cudaFuncSetAttribute, occupancy and the SM count are PER DEVICE: every "already done" flag / cached value in the
// launchers is an array indexed by the current device (a process may drive several GPUs through one library).
inline int device_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= NF_MAX_DEVICES) d = 0;
    return d;
}
// thread-local error string behind nf_last_error() (defined in nf_api.cu); returns `code`
int set_error(int code, const char* what, const char* msg);
cudaError_t launch_chain(const NfModelParams& mp, const NfChainArgs& args, bool inverse, int num_sms, int warps_per_cta,
                         cudaStream_t stream);
cudaError_t launch_reduce(const float* nll, const float* sdz, long long n, double* sums, cudaStream_t stream);
cudaError_t launch_squeeze(const float* in, float* out, long long n, int H, int W, int C, int factor, int patch_type,
                           int inverse, cudaStream_t stream);
bool program_is_scale_only(const NfModelParams& mp, int first, int last);
cudaError_t launch_scale_stream(const NfModelParams& mp, const NfChainArgs& args, bool inverse, int num_sms,
                                cudaStream_t stream);
cudaError_t launch_baseline_nll(const float* x, const float* y, float nlf0, float nlf1, float var_gauss, long long n,
                                float* nll_gauss, float* nll_sdn, int num_sms, cudaStream_t stream);
cudaError_t launch_histogram(const float* data, long long count, const double* edges, int n_bins, unsigned long long* counts,
                             int num_sms, cudaStream_t stream);
bool tc_program_supported(const NfModelParams& mp, const NfChainArgs& a);
cudaError_t launch_chain_tc(const NfModelParams& mp, const NfChainArgs& args, bool inverse, int num_sms, cudaStream_t stream);
// hybrid chain kernel (nf_hybrid.cu): conv-3 of every coupling on tcgen05, everything else as in launch_chain
bool hybrid_program_supported(const NfModelParams& mp, const NfChainArgs& a);
cudaError_t launch_chain_hybrid(const NfModelParams& mp, const NfChainArgs& args, bool inverse, int num_sms, cudaStream_t stream);
// all-fp32 chain kernel with vertical Winograd F(2,3) 3x3 convolutions (nf_wino.cu)
bool wino_program_supported(const NfModelParams& mp, const NfChainArgs& a);
cudaError_t launch_chain_wino(const NfModelParams& mp, const NfChainArgs& args, bool inverse, int num_sms, cudaStream_t stream);
}  // namespace nf
