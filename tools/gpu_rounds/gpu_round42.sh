#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zb.json 2> gpurun_out/bench_zb.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_zb_l1.json 2> gpurun_out/bench_zb_l1.err
MTL_ZSLAB_CTAS=296 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_zb_l1_296.json 2> gpurun_out/bench_zb_l1_296.err
MTL_ZSLAB_CTAS=80 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_zb_l1_80.json 2> gpurun_out/bench_zb_l1_80.err
echo done
