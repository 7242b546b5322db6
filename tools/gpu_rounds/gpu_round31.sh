#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err
MTL_PRIORITIES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r_noprio.json 2> gpurun_out/bench_r_noprio.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_r_l1.json 2> gpurun_out/bench_r_l1.err
echo done
