#!/bin/bash
# parity subset, then the default chain kernel with and without its L2 prefetch (NF_WINO_PREFETCH=0):  bash tools/gpu/ab_prefetch.sh
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_winograd.py tests/test_gpu_reference_goldens.py -x -q -m gpu 2>&1 | tail -4
for pf in 1 0 1 0; do echo "NF_WINO_PREFETCH=$pf"; NF_WINO_PREFETCH=$pf bash tools/gpu/ab_quick.sh; done 2>&1 | tee gpurun_out/ab_prefetch.log
