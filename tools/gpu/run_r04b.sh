#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu/wide_tc_check.py 32 64 128 2>&1 | tee gpurun_out/r04b_wide_tc_check.log
