#!/bin/bash
# round-3k: final single-GPU verification: whole suite, smoke, sanitizer (4 tools), benches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_full.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
NF_SANITIZE_TOOLS="memcheck racecheck initcheck" bash tools/gpu/scripts_gpu_sanitize.sh 2>&1 | grep -E "^==|SUMMARY|cooperative" | tee gpurun_out/sanitize_summary.log
NF_SANITIZE_CHAIN=0 NF_SANITIZE_N=37 timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool synccheck --print-limit 20 python tools/gpu/sanitize_smoke.py > gpurun_out/sanitize_synccheck.log 2>&1
echo "== synccheck (NF_SANITIZE_CHAIN=0) rc=$?" | tee -a gpurun_out/sanitize_summary.log; grep -E "ERROR SUMMARY|^ok|cooperative" gpurun_out/sanitize_synccheck.log | tee -a gpurun_out/sanitize_summary.log
timeout 400 python bench.py > gpurun_out/r03_bench_default.json 2> gpurun_out/r03_bench_default.err
timeout 300 python bench.py --mode sample --no-cpu-baseline > gpurun_out/r03_bench_sample.json 2> gpurun_out/r03_bench_sample.err
rm -f gpurun_out/r03_bench_train.jsonl
for b in 138 207 2048; do timeout 200 python bench.py --mode train --batch $b --steps 200 --no-cpu-baseline >> gpurun_out/r03_bench_train.jsonl 2>>gpurun_out/r03_bench_default.err; done
timeout 300 python tools/gpu/wrapper_latency.py 2>&1 | tail -10 > gpurun_out/r03_wrapper_latency.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03_launches_train_138_fused.csv python bench.py --mode train --batch 138 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_train.log 2>&1
python - <<'PY'
import json
for f in ('r03_bench_default','r03_bench_sample'):
    try:
        j=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, '%.3fM/s' % (j['value']/1e6), 'e2e', j['e2e'] and '%.3fM/s' % (j['e2e']['value']/1e6), (j.get('roofline') or {}).get('frac'), (j.get('roofline_fp32') or {}).get('frac'))
    except Exception as e: print(f, 'FAILED', e)
for l in open('gpurun_out/r03_bench_train.jsonl'):
    j=json.loads(l); print('train batch', j['config']['per_gpu_batch'], '%.3f ms/step' % j['ms_per_step'], '%.0f patches/s' % j['value'], 'e2e', j['e2e'] and round(j['e2e']['value']), 'launches', j['gpu_launches'])
PY
head -4 gpurun_out/r03_wrapper_latency.txt
tail -q -n 3 gpurun_out/r03_bench_default.err gpurun_out/r03_bench_sample.err
