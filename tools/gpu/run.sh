#!/bin/bash
# One entry point for everything that runs on the GPU box:   gpurun --timeout 1500 -- 'bash tools/gpu/run.sh <task> [args]'
#   suite            pytest -m gpu (whole suite)                      tests <pytest args>   a subset, e.g. tests/test_gpu_wide.py -k gradients
#   bench [args]     python bench.py [args] -> gpurun_out/bench.json + a summary of the `also` block
#   wide             bench.py --width 32 / 64 / 128 / 256 / 512        widecheck [widths]    tensor-core wide kernels vs the fp64 oracle
#   multi N          bench at N GPUs under torchrun (also-block, shard_check, e2e ceiling) + data-parallel train checks
#   ncu-chain        ncu --set full of the shipped chain kernels      ncu-wide W            ... of the tensor-core wide kernel at width W
#   launches [args]  ncu launch list of bench.py [args]               probe                 tools/tc_probe2.cu + PCIe probe
#   sanitize         compute-sanitizer smoke (tools/gpu/sanitize_smoke.py)
# Round 5 companions: tools/gpu/ab_chain.sh (A/B of the three width-4 chain kernels), sanitize_chain.sh (sanitizer over them),
# final_evidence.sh (suite + default bench + ncu --set full of the shipped chain kernels + launch list in one call).
# Round 6: ab_quick.sh (default chain kernel, both directions, 20 steps), ab_cta_warps.sh (one 16-warp CTA per SM vs two 8-warp CTAs).
set -u
mkdir -p gpurun_out
task=${1:-suite}; shift || true
summ() { python - "$1" <<'PY'
import json, sys
j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
e = j.get('e2e') or {}
print('N=%d value %.3f M  e2e %.3f M  roofline %s frac %.4f' % (j['n_gpus'], j['value'] / 1e6, (e.get('value') or 0) / 1e6, (j.get('roofline') or {}).get('bound'), (j.get('roofline') or {}).get('frac') or -1))
if e.get('ceiling_gbs'): print('  e2e h2d %.1f GB/s of ceiling %.1f = %.2f' % (e['h2d_gbs'], e['ceiling_gbs'], e['frac_of_ceiling']))
if j.get('shard_check'): print('  shard_check', j['shard_check'])
for k, v in (j.get('also') or {}).items():
    r = v['roofline']
    print('  %-32s %10.3f k patches/s  %.3f ms  %s frac %.3f' % (k, v['value'] / 1e3, v['ms_per_step'], r['bound'], r['frac'] or -1))
PY
}
case $task in
  suite) timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log ;;
  tests) timeout 1500 python -m pytest "$@" -q -m gpu -s 2>&1 | grep -vE "^\s*$" | tail -60 | tee gpurun_out/pytest_subset.log ;;
  bench) timeout 900 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc $?"; tail -3 gpurun_out/bench.err; summ gpurun_out/bench.json ;;
  wide) for w in 32 64 128 256 512; do b=16384; [ $w -ge 256 ] && b=4096
          timeout 300 python bench.py --width $w --batch $b --steps 10 --no-cpu-baseline --no-e2e 2>>gpurun_out/wide.err; done | tee gpurun_out/bench_wide.jsonl | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print('width', j['config']['width'], '%.3f ms' % j['ms_per_step'], '%.3f M patches/s' % (j['value'] / 1e6), 'issued frac %.3f' % j['roofline']['issued_frac'])" ;;
  widecheck) timeout 600 python tools/gpu/wide_tc_check.py "$@" 2>&1 | tee gpurun_out/wide_tc_check.log ;;
  multi) N=${1:-2}
     timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
     echo "bench rc $?"; grep -vE '^\*|OMP_NUM' gpurun_out/bench_n$N.err | tail -5; summ gpurun_out/bench_n$N.json
     timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/gpu/dp_train_check.py 2>&1 | grep -vE '^\*|OMP_NUM' | tail -8 | tee gpurun_out/dp_train_check_n$N.log ;;
  ncu-chain) for m in log_prob sample; do
       timeout 400 ncu --set full --clock-control none --import-source on -k regex:nf_chain_kernel -s 3 -c 1 -o gpurun_out/prof_chain_$m -f python bench.py --mode $m --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also > gpurun_out/ncu_chain_$m.log 2>&1; done; ls -la gpurun_out/*.ncu-rep ;;
  ncu-wide) W=${1:-32}; b=16384; [ $W -ge 256 ] && b=1024
     timeout 400 ncu --set full --clock-control none --import-source on -k regex:nf_wide_tc -s 3 -c 1 -o gpurun_out/prof_wide_tc$W -f python bench.py --width $W --batch $b --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_wide_tc$W.log 2>&1; ls -la gpurun_out/*.ncu-rep ;;
  launches) timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-also "$@" > gpurun_out/ncu_launches.log 2>&1
     grep -E "nf_|td_" gpurun_out/launches.csv | tail -12 | cut -d, -f5,12- | cut -c1-160 ;;
  probe) nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tc_probe2 tools/tc_probe2.cu && timeout 120 gpurun_out/tc_probe2 2>&1 | tee gpurun_out/tc_probe2.log
     timeout 120 python tools/gpu/pcie_probe.py 2>&1 | tee gpurun_out/pcie_probe.log ;;
  sanitize) for tool in memcheck racecheck initcheck synccheck; do echo "== $tool"; c=1; [ $tool = synccheck ] && c=0   # synccheck (CUDA 12.9) trips over tcgen05.alloc of the TMEM-resident chain kernel
       NF_SANITIZE_CHAIN=$c timeout 900 compute-sanitizer --tool $tool python tools/gpu/sanitize_smoke.py 2>&1 | tail -4; done | tee gpurun_out/compute_sanitizer.txt ;;
  *) echo "unknown task $task"; exit 2 ;;
esac
