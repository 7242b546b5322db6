#!/bin/bash
set -u
mkdir -p gpurun_out
MTL_CONV_WGRAD_2CTA=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k conv3x3 > gpurun_out/pytest_conv2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_conv2.log
MTL_CONV_WGRAD_2CTA=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zf.json 2> gpurun_out/bench_zf.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zf0.json 2> gpurun_out/bench_zf0.err
echo done
