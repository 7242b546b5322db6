#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo done
