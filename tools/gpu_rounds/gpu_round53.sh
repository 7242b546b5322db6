#!/bin/bash
set -u
mkdir -p gpurun_out
MTL_CONV_WGRAD_KW=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "conv3x3" > gpurun_out/pytest_wkw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_wkw.log
{ MTL_CONV_WGRAD_KW=1 timeout 60 python tools/probes/one_conv_bwd.py 2 20 3; timeout 60 python tools/probes/one_conv_bwd.py 2 20 3;
  MTL_CONV_WGRAD_KW=1 timeout 60 python tools/probes/one_conv_bwd.py 1 20 3; timeout 60 python tools/probes/one_conv_bwd.py 1 20 3; } > gpurun_out/conv_wkw.log 2>&1
echo done
