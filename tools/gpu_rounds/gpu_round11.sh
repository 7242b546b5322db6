#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_f_l1.json 2> gpurun_out/bench_f_l1.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 --gemm-mode 1 > gpurun_out/bench_f_l1_tf32.json 2> gpurun_out/bench_f_l1_tf32.err
MTL_BRANCHES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_f_l1_br0.json 2> gpurun_out/bench_f_l1_br0.err
echo done
