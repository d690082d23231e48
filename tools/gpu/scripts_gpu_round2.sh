#!/bin/bash
# round-2 GPU visit: whole GPU suite, wide-net throughput, train-step launch list, ncu --set full of the new kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
rm -f gpurun_out/wide.jsonl gpurun_out/wide.err
for w in 8 16 32; do timeout 300 python bench.py --width $w --batch 16384 --steps 10 --no-cpu-baseline --no-e2e >> gpurun_out/wide.jsonl 2>>gpurun_out/wide.err; done
python - <<'PY'
import json
for l in open('gpurun_out/wide.jsonl'):
    j=json.loads(l); print('width', j['config'].get('width'), '%.3f ms/step' % j['ms_per_step'], '%.3f M patches/s' % (j['value']/1e6), 'fp32 frac %.3f' % j['roofline_fp32']['frac'])
PY
tail -3 gpurun_out/wide.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_train_207.csv python bench.py --mode train --batch 207 --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_launch_train.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:td_b1_kernel -s 20 -c 1 -o gpurun_out/prof_td_b1 -f python bench.py --mode train --batch 207 --steps 2 --warmup 3 --no-e2e > gpurun_out/ncu_full_b1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nf_wide_chain -s 3 -c 1 -o gpurun_out/prof_wide32 -f python bench.py --width 32 --batch 4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_wide.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_launch_train.log
