#!/bin/bash
# usage: run_multi.sh N  -- bench at N GPUs (also-block, shard_check, e2e ceiling), data-parallel train check
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/r04_bench_n$N.json 2> gpurun_out/r04_bench_n$N.err
echo "bench rc $?"; tail -8 gpurun_out/r04_bench_n$N.err
python - <<PY
import json
j = json.loads([l for l in open('gpurun_out/r04_bench_n$N.json').read().splitlines() if l.startswith('{')][-1])
e = j['e2e']
print('N=%d value %.3f M  e2e %.3f M (h2d %.1f GB/s of ceiling %.1f = %.2f)' % (j['n_gpus'], j['value']/1e6, e['value']/1e6, e['h2d_gbs'], e['ceiling_gbs'], e['frac_of_ceiling']))
print('shard_check', j.get('shard_check'))
for k, v in (j.get('also') or {}).items():
    print('  %-32s %10.3f k patches/s  %.3f ms' % (k, v['value']/1e3, v['ms_per_step']))
PY
if [ -f tools/gpu/dp_train_check.py ]; then timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/gpu/dp_train_check.py 2>&1 | tail -6 | tee gpurun_out/r04_dp_train_check_n$N.log; fi
