#!/bin/bash
set -u
mkdir -p gpurun_out
MTL_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zh_nopdl.json 2> gpurun_out/bench_zh_nopdl.err
MTL_ATTN_THREADS=256 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zh_a256.json 2> gpurun_out/bench_zh_a256.err
MTL_WGRAD_CTAS=12 MTL_DGRAD_CTAS=12 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zh_12.json 2> gpurun_out/bench_zh_12.err
echo done
