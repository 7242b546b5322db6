#!/bin/bash
set -u
mkdir -p gpurun_out
{ for m in 2 1; do
  MTL_GEMM_DBG=99 timeout 120 python tests/gpu_gemm_latency.py $m
  timeout 120 python tests/gpu_gemm_latency.py $m
  MTL_CLUSTER_SPLITK=0 timeout 120 python tests/gpu_gemm_latency.py $m
done; } > gpurun_out/gemm_latency.log 2>&1
echo done
