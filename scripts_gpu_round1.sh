#!/bin/bash
# one GPU-box visit: parity tests, bench, sweep, ncu launch list + full capture of the chain kernel
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --mode sample --no-cpu-baseline --steps 10 > gpurun_out/bench_sample.json 2>gpurun_out/bench_sample.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nf_chain_kernel -s 3 -c 1 -o gpurun_out/prof_chain -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
python - <<'PY'
import json
for f in ('bench_default','bench_sample'):
    j=json.load(open('gpurun_out/%s.json'%f)); print(f, '%.3fM/s' % (j['value']/1e6), 'kernel_ms %.3f' % j['roofline']['kernel_ms'], 'hbm frac %.4f' % j['roofline']['frac'], 'fp32frac %.3f' % j['roofline_fp32']['frac'], 'e2e %.3fM/s' % (j['e2e']['value']/1e6), j['clocks'])
PY
