#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv3x3" > gpurun_out/pytest_conv_kw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_conv_kw.log
{ for m in 2 1; do for l in 1 2 3; do timeout 120 python tools/probes/one_conv.py $m 20 $l; MTL_CONV_KW=0 timeout 120 python tools/probes/one_conv.py $m 20 $l; done; done; } > gpurun_out/conv_kw_ab.log 2>&1
echo done
