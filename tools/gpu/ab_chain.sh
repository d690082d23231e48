#!/bin/bash
# A/B of the width-4 chain kernels on one GPU:  bash tools/gpu/ab_chain.sh [modes...]   (--tc 5 = direct-form all-fp32, 4 = Winograd all-fp32, 2 = hybrid)
mkdir -p gpurun_out
: > gpurun_out/ab_chain.jsonl
for mode in log_prob sample; do for tc in 5 4 2 ${@}; do   # explicit kernels, not the library default
  timeout 150 python bench.py --mode $mode --tc $tc --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-also 2>>gpurun_out/ab_chain.err | tail -1 >> gpurun_out/ab_chain.jsonl
done; done
python - <<'PY'
import json
for l in open('gpurun_out/ab_chain.jsonl'):
    if not l.startswith('{'): continue
    j = json.loads(l)
    print(j['config']['workload'][:10], '%.3f M patches/s' % (j['value'] / 1e6), '%.3f ms' % j['ms_per_step'], 'nll/dim', j.get('mean_nll_per_dim'))
PY
