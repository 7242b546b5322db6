#!/bin/bash
set -u
mkdir -p gpurun_out
{ timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k cfg4 2>&1 | grep -E "cfg4 worst|passed|failed"
  echo "--- RN split"; MTL_SPLIT_TRUNC=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k cfg4 2>&1 | grep -E "cfg4 worst|passed|failed"
  echo "--- fp32 SIMT"; MTL_GEMM_MODE=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k cfg4 2>&1 | grep -E "cfg4 worst|passed|failed"
} > gpurun_out/cfg4.log 2>&1
echo done
