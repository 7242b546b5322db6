#!/bin/bash
set -u
mkdir -p gpurun_out
{ MTL_GEMM_DBG=1 python tools/probes/one_conv.py 2 6 1 | tail -17;  MTL_EPI_TEST=1 MTL_GEMM_DBG=1 python tools/probes/one_conv.py 2 6 1 | tail -17; MTL_GEMM_DBG=1 python tools/probes/one_conv.py 1 20 1 | tail -17; } > gpurun_out/conv_stamps2.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/conv_stamps2.log
echo done
