#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "conv3x3" > gpurun_out/pytest_conv_kw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_conv_kw.log
echo done
