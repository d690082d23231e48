"""Facts about the compiled kernels that host code needs (kept next to the sources they describe)."""
WIDE_CUDA_CORE_WIDTHS = (8, 16, 32)            # csrc/nf_wide.cu
WIDE_TC_WIDTHS = (32, 64, 128, 256, 512)       # csrc/nf_wide_tc.cu (resident weights: 32 / 64 / 128), nf_wide_tcs.cu (streamed: 256 / 512)
