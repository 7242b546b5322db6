#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 296 96 48 24; do
  MTL_WGRAD_CTAS=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t_$v.json 2> gpurun_out/bench_t_$v.err
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_t_l1.json 2> gpurun_out/bench_t_l1.err
echo done
