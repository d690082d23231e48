#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "batch_stat or training_mode or threads" 2>&1 | tail -3
for i in 1 2; do timeout 300 python tools/gpu/wrapper_latency.py 2>&1 | grep "bn_mode=batch" | head -3; done | tee gpurun_out/r03_wrapper_latency_b.txt
