#!/bin/bash
# A/B of environment switches: `bash tools/ab_bench.sh "VAR=1 VAR2=2" "VAR=3" ...` -> one line per configuration
# (ms/step device-resident, e2e ms/step) in gpurun_out/ab_bench.log
mkdir -p gpurun_out
for cfg in "$@"; do
  line=$(env $cfg timeout 300 python bench.py --no-cpu-baseline --no-gpu-baseline --no-roofline --steps 10 --warmup 3 2>gpurun_out/ab_err.log | tail -1)
  ms=$(echo "$line" | python -c 'import json,sys
try:
    d=json.loads(sys.stdin.read()); print("%.3f ms/step  e2e %.3f  launches/step %d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"]/d["steps"]))
except Exception as e:
    print("FAILED", e)')
  echo "$cfg :: $ms" | tee -a gpurun_out/ab_bench.log
  [ "${ms:0:6}" = "FAILED" ] && tail -5 gpurun_out/ab_err.log
done
