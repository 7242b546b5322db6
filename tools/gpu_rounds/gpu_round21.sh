#!/bin/bash
set -u
mkdir -p gpurun_out
{ timeout 120 python tests/gpu_gemm_latency.py 2
  MTL_PDL=0 timeout 120 python tests/gpu_gemm_latency.py 2
  timeout 120 python tests/gpu_gemm_latency.py 1
} > gpurun_out/gemm_latency2.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err
MTL_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_i_nopdl.json 2> gpurun_out/bench_i_nopdl.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_i_l1.json 2> gpurun_out/bench_i_l1.err
echo done
