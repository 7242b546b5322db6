#!/bin/bash
set -u
mkdir -p gpurun_out
MTL_DGRAD_CTAS=12 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w_a.json 2> gpurun_out/bench_w_a.err
MTL_DGRAD_CTAS=24 MTL_CLUSTER_MAX=4 MTL_CLUSTER_CTAS=32 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w_b.json 2> gpurun_out/bench_w_b.err
MTL_DGRAD_CTAS=24 MTL_CLUSTER_MAX=2 MTL_CLUSTER_CTAS=64 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w_c.json 2> gpurun_out/bench_w_c.err
MTL_DGRAD_CTAS=24 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_w_d.json 2> gpurun_out/bench_w_d.err
MTL_DGRAD_CTAS=24 MTL_WGRAD_CTAS=24 MTL_CONV_WGRAD_WAVES=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w_e.json 2> gpurun_out/bench_w_e.err
echo done
