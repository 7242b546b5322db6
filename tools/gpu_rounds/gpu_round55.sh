#!/bin/bash
set -u
mkdir -p gpurun_out
MTL_CONV_WGRAD_KW64=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "conv3x3" > gpurun_out/pytest_wkw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_wkw.log
{ for l in 1 2; do MTL_CONV_WGRAD_KW64=1 timeout 60 python tools/probes/one_conv_bwd.py 2 20 $l; timeout 60 python tools/probes/one_conv_bwd.py 2 20 $l; done; } > gpurun_out/conv_wkw64.log 2>&1
echo done
