#!/bin/bash
# runs the tcgen05 GEMM probe over layouts/shapes, one process per case
cd "$(dirname "$0")/.."
for dt in 0; do
  export MTL_TMA_FP32=$dt
  echo "== MTL_TMA_FP32=$dt"
  for lay in "0 0" "1 0" "1 1"; do
    for shp in "128 128 32" "128 64 64" "264 100 512" "1000 64 576"; do
      timeout 60 python tests/gpu_gemm_probe.py $lay $shp 2>&1 | tail -2
    done
  done
done
export MTL_TMA_FP32=0
timeout 60 python tests/gpu_gemm_probe.py 1 0 64 576 4000 8 2>&1 | tail -2
timeout 60 python tests/gpu_gemm_probe.py 1 0 128 1152 32000 16 2>&1 | tail -2
