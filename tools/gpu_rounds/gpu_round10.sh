#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
{ for m in 2; do for l in 1 2 3; do python tools/probes/one_conv.py $m 20 $l; done; done
  MTL_GEMM_DBG=600 python tools/probes/one_conv.py 2 6 1 | tail -22; } > gpurun_out/conv_ab.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err
echo done
