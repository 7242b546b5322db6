#!/bin/bash
set -u
mkdir -p gpurun_out
{ for m in 2 1; do
    echo "== mode $m slab wgrad 512x512x264 split 9"; MTL_GEMM_DBG=1 timeout 60 python tests/gpu_one_gemm.py $m 1 0 512 512 264 4 1.0 9
    echo "== mode $m cluster 264x512x100"; MTL_GEMM_DBG=1 timeout 60 python tests/gpu_one_gemm.py $m 0 1 264 512 100 4
    echo "== mode $m cluster 264x100x512"; MTL_GEMM_DBG=1 timeout 60 python tests/gpu_one_gemm.py $m 0 1 264 100 512 4
done; } > gpurun_out/gemm_stamps2.log 2>&1
echo done
