#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for m in 2 1; do for l in 1 2 3; do python tools/probes/one_conv.py $m 20 $l; done; done > gpurun_out/conv_ab.log 2>&1
for l in 1; do MTL_CONV_STAGES=2 python tools/probes/one_conv.py 2 20 $l; done >> gpurun_out/conv_ab.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_br1.json 2> gpurun_out/bench_br1.err
MTL_BRANCHES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_br0.json 2> gpurun_out/bench_br0.err
MTL_CONV_STAGES=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_br1_st2.json 2> gpurun_out/bench_br1_st2.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-mode 1 > gpurun_out/bench_br1_tf32.json 2> gpurun_out/bench_br1_tf32.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_br1_l1.json 2> gpurun_out/bench_br1_l1.err
echo done
