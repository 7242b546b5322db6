#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_o.json 2> gpurun_out/bench_o.err
MTL_SPLIT_TRUNC=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_o_rn.json 2> gpurun_out/bench_o_rn.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_o_l1.json 2> gpurun_out/bench_o_l1.err
{ python tools/probes/one_conv.py 2 20 1; python tools/probes/one_conv.py 2 20 3; MTL_SPLIT_TRUNC=0 python tools/probes/one_conv.py 2 20 1; } > gpurun_out/conv_trunc.log 2>&1
echo done
