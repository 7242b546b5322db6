#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 96 48 24; do
  MTL_DGRAD_CTAS=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v_$v.json 2> gpurun_out/bench_v_$v.err
done
echo done
