#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 --timeline gpurun_out/timeline_y_1l.csv.gz > gpurun_out/bench_y_l1.json 2> gpurun_out/bench_y_l1.err
echo done
