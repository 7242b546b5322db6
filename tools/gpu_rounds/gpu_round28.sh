#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tests/gpu_gemm_latency.py 2 > gpurun_out/gemm_latency3.log 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "gemm" > gpurun_out/pytest_gemm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gemm.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err
echo done
