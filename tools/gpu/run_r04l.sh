#!/bin/bash
mkdir -p gpurun_out
export TORCH_DISTRIBUTED_DEBUG=OFF
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 --no-also > gpurun_out/l1.json 2> gpurun_out/l1.err; echo "rc noalso $?"; tail -c 600 gpurun_out/l1.json; tail -5 gpurun_out/l1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/l2.json 2> gpurun_out/l2.err; echo "rc also $?"; tail -c 300 gpurun_out/l2.json; tail -25 gpurun_out/l2.err
