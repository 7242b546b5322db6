#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-mode 1 > gpurun_out/bench_z_tf32.json 2> gpurun_out/bench_z_tf32.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_z_l1.json 2> gpurun_out/bench_z_l1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_z_ref.json 2> gpurun_out/bench_z_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo done
