#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
{ python tools/probes/attn_probe.py; echo "--- 256 threads"; MTL_ATTN_THREADS=256 python tools/probes/attn_probe.py; } > gpurun_out/attn_probe.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_l_l1.json 2> gpurun_out/bench_l_l1.err
echo done
