#!/bin/bash
# one GPU-box visit: parity tests, bench (all arms), ncu launch list + full capture of the chain kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 200 python bench.py --mode sample --no-cpu-baseline > gpurun_out/bench_sample.json 2>gpurun_out/bench_sample.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err
timeout 200 python bench.py --arch "sdn5|gain4" --no-cpu-baseline --no-e2e > gpurun_out/bench_sdn_gain.json 2> gpurun_out/bench_sdn_gain.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:nf_chain_kernel -s 3 -c 1 -o gpurun_out/prof_chain -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:nf_scale_stream -s 3 -c 1 -o gpurun_out/prof_stream -f python bench.py --arch "sdn5|gain4" --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_stream.log 2>&1
python - <<'PY'
import json
for f in ('bench_default','bench_sample','bench_sdn_gain'):
    try:
        j=json.load(open('gpurun_out/%s.json'%f)); print(f, '%.3fM/s' % (j['value']/1e6), 'kernel_ms %.3f' % j['roofline']['kernel_ms'], 'hbm frac %.4f' % j['roofline']['frac'], 'fp32frac %.3f' % j['roofline_fp32']['frac'], 'e2e', j['e2e'] and '%.3fM/s' % (j['e2e']['value']/1e6), j['clocks'])
    except Exception as e: print(f, 'FAILED', e)
PY
