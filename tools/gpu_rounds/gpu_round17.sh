#!/bin/bash
set -u
mkdir -p gpurun_out
{ for sel in 1 0; do echo "SEL $sel"; MTL_CONV_KW_SEL=$sel timeout 120 python tools/probes/c1_probe.py; done; } > gpurun_out/c1_probe.log 2>&1
echo done
