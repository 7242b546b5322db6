#!/bin/bash
set -u
mkdir -p gpurun_out
MTL_CONV_WGRAD_BN64=1 MTL_CONV_KW_BN64=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k conv3x3 > gpurun_out/pytest_conv2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_conv2.log
MTL_CONV_WGRAD_BN64=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zg_w.json 2> gpurun_out/bench_zg_w.err
MTL_CONV_KW_BN64=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zg_k.json 2> gpurun_out/bench_zg_k.err
MTL_CONV_WGRAD_BN64=1 MTL_CONV_KW_BN64=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zg_wk.json 2> gpurun_out/bench_zg_wk.err
echo done
