#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/probes/attn_probe.py > gpurun_out/attn_probe.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err
MTL_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m_nopdl.json 2> gpurun_out/bench_m_nopdl.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_m_l1.json 2> gpurun_out/bench_m_l1.err
MTL_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_m_l1_nopdl.json 2> gpurun_out/bench_m_l1_nopdl.err
echo done
