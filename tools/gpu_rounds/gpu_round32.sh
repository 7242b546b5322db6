#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "4 160" "2 160" "8 64" "4 64"; do set -- $v
  MTL_CLUSTER_MAX=$1 MTL_CLUSTER_CTAS=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s_$1_$2.json 2> gpurun_out/bench_s_$1_$2.err
done
echo done
