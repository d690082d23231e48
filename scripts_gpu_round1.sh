#!/bin/bash
# one GPU-box visit: parity tests, bench, sweep, ncu launch list + full capture of the chain kernel
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for w in 4 6 8 10 12; do python bench.py --warps $w --no-cpu-baseline --no-e2e --steps 10 >> gpurun_out/sweep.jsonl 2>>gpurun_out/sweep.err; done
python bench.py --batch 4096 --no-cpu-baseline --no-e2e --steps 50 >> gpurun_out/sweep.jsonl 2>>gpurun_out/sweep.err
python bench.py --mode sample --no-cpu-baseline --steps 10 > gpurun_out/bench_sample.json 2>gpurun_out/bench_sample.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nf_chain_kernel -s 3 -c 1 -o gpurun_out/prof_chain -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
