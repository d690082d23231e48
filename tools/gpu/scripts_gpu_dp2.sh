#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/gpu/dp_train_check.py 2>&1 | grep -v "^W\|^\*\|OMP_NUM" | tail -8 | tee gpurun_out/r02_dp_train_check_n2.log
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --mode train --batch 138 --steps 100 > gpurun_out/r02_bench_train_n2.json 2> gpurun_out/bench_train_n2.err
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 30 --no-cpu-baseline > gpurun_out/r02_bench_n2.json 2> gpurun_out/bench_n2.err
cat gpurun_out/r02_bench_train_n2.json | cut -c1-400; tail -3 gpurun_out/bench_train_n2.err
cat gpurun_out/r02_bench_n2.json | cut -c1-300; tail -3 gpurun_out/bench_n2.err
