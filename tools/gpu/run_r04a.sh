#!/bin/bash
# round-4a: TMEM-A MMA probe, refreshed ncu --set full of the SHIPPED chain kernels (both directions), PCIe probe
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tc_probe2 tools/tc_probe2.cu
timeout 120 gpurun_out/tc_probe2 2>&1 | tee gpurun_out/r04_tc_probe2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nf_chain_kernel -s 3 -c 1 -o gpurun_out/r04_prof_chain_inv -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_chain_inv.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nf_chain_kernel -s 3 -c 1 -o gpurun_out/r04_prof_chain_fwd -f python bench.py --mode sample --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_chain_fwd.log 2>&1
timeout 120 python tools/gpu/pcie_probe.py 2>&1 | tee gpurun_out/r04_pcie_probe_n1.log
ls -la gpurun_out/*.ncu-rep
