#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-mode 1 > gpurun_out/bench_d_tf32.json 2> gpurun_out/bench_d_tf32.err
echo done
