#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nf_wide_tc_kernel -s 3 -c 1 -o gpurun_out/r04_prof_wide_tc32 -f python bench.py --width 32 --batch 16384 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_wide_tc32.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nf_wide_tcs_kernel -s 3 -c 1 -o gpurun_out/r04_prof_wide_tcs512 -f python bench.py --width 512 --batch 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_wide_tcs512.log 2>&1
for w in 32 64 128 256 512; do b=16384; [ $w -ge 256 ] && b=4096; timeout 300 python bench.py --width $w --batch $b --steps 10 --no-cpu-baseline --no-e2e 2>>gpurun_out/r04e_wide.err; done | tee gpurun_out/r04e_bench_wide.jsonl | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print('width', j['config']['width'], '%.3f ms' % j['ms_per_step'], '%.3f M patches/s' % (j['value'] / 1e6))
"
tail -3 gpurun_out/r04e_wide.err; ls -la gpurun_out/*.ncu-rep
