#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tc_probe tools/tc_probe.cu
timeout 60 gpurun_out/tc_probe 2>&1 | tee gpurun_out/tc_probe.log
