#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 12 6 1; do
  MTL_WGRAD_CTAS=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t_$v.json 2> gpurun_out/bench_t_$v.err
done
MTL_WGRAD_CTAS=24 MTL_CLUSTER_MAX=8 MTL_CLUSTER_CTAS=160 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_t_l1_24.json 2> gpurun_out/bench_t_l1_24.err
MTL_WGRAD_CTAS=6 MTL_CLUSTER_MAX=8 MTL_CLUSTER_CTAS=160 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_t_l1_6.json 2> gpurun_out/bench_t_l1_6.err
echo done
