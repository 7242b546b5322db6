#!/bin/bash
set -u
mkdir -p gpurun_out
{ MTL_GEMM_DBG=1 python tools/probes/one_conv.py 2 6 1; MTL_GEMM_DBG=600 python tools/probes/one_conv.py 2 6 1; MTL_GEMM_DBG=600 python tools/probes/one_conv.py 1 6 1; MTL_GEMM_DBG=600 MTL_CONV_STAGES=2 python tools/probes/one_conv.py 2 6 1; } > gpurun_out/conv_stamps.log 2>&1
echo done
