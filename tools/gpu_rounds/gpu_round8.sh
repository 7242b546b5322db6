#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 3200 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo done
