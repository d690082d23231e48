#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err

python - <<PY
import json
j=json.load(open('gpurun_out/bench_n$N.json')); print('N=$N', '%.3fM/s' % (j['value']/1e6), 'ms/step %.3f' % j['ms_per_step'], 'e2e %.3fM/s' % (j['e2e']['value']/1e6), j['config']['mean_nll_per_dim'], j['clocks'])

PY
