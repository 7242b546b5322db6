#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
{ for m in 2 1; do for l in 1 2 3; do python tools/probes/one_conv.py $m 20 $l; done; done
  MTL_CONV_STAGES=2 python tools/probes/one_conv.py 2 20 1
  MTL_GEMM_DBG=600 python tools/probes/one_conv.py 2 6 1
  MTL_GEMM_DBG=1 python tools/probes/one_conv.py 2 6 1 | tail -9; } > gpurun_out/conv_ab.log 2>&1
python tests/gpu_microbench.py > gpurun_out/microbench.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
MTL_CONV_STAGES=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a_st2.json 2> gpurun_out/bench_a_st2.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-mode 1 > gpurun_out/bench_a_tf32.json 2> gpurun_out/bench_a_tf32.err
echo done
