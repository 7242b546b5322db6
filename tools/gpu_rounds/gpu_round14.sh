#!/bin/bash
set -u
mkdir -p gpurun_out
{ for w in both b1 b2; do
  timeout 120 python tools/probes/accum_probe.py $w
  MTL_CONV_KW=0 timeout 120 python tools/probes/accum_probe.py $w
  MTL_BRANCHES=0 timeout 120 python tools/probes/accum_probe.py $w
done; } > gpurun_out/accum_probe.log 2>&1
echo done
