#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tc_probe tools/tc_probe.cu
timeout 60 gpurun_out/tc_probe 2>&1 | tee gpurun_out/tc_probe.log
python bench.py --arch "sdn5|gain4" --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/bench_sdn_gain.json 2> gpurun_out/bench_sdn_gain.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_sdn_gain.json')); print('sdn5|gain4', '%.2fM/s' % (j['value']/1e6), 'kernel_ms %.3f' % j['roofline']['kernel_ms'], 'hbm GB/s %.0f frac %.3f' % (j['roofline']['achieved'], j['roofline']['frac']))
PY
