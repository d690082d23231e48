#!/bin/bash
# round-3: two-GPU checks: data-parallel train replicas stay identical (fused step kernel + one all-reduce), benches, reference arm
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/gpu/dp_train_check.py 2>&1 | grep -v "^W\|^\*\|OMP_NUM" | tail -8 | tee gpurun_out/r03_dp_train_check_n2.log
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --mode train --batch 138 --steps 200 --no-cpu-baseline > gpurun_out/r03_bench_train_n2.json 2> gpurun_out/bench_train_n2.err
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 50 --no-cpu-baseline > gpurun_out/r03_bench_n2.json 2> gpurun_out/bench_n2.err
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --mode sample --steps 50 --no-cpu-baseline > gpurun_out/r03_bench_sample_n2.json 2> gpurun_out/bench_sample_n2.err
timeout 200 $TR --master-port 29515 bench.py --impl reference --gpus 2 --steps 5 > gpurun_out/r03_bench_reference_n2.json 2> gpurun_out/bench_ref_n2.err
for f in r03_bench_train_n2 r03_bench_n2 r03_bench_sample_n2 r03_bench_reference_n2; do python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', j['metric'], '%.4gM/s' % (j['value']/1e6), 'ms/step %.3f' % j['ms_per_step'], 'e2e', j['e2e'] and '%.4gM/s' % (j['e2e']['value']/1e6), 'n_gpus', j['n_gpus'])
except Exception as e: print('$f FAILED', e)
PY
done
tail -q -n 2 gpurun_out/bench_train_n2.err gpurun_out/bench_n2.err gpurun_out/bench_sample_n2.err gpurun_out/bench_ref_n2.err | grep -v "^W\|^\*\|OMP_NUM" | head
