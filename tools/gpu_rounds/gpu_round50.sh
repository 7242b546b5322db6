#!/bin/bash
set -u
mkdir -p gpurun_out
{ for l in 1 2 3; do python tools/probes/one_conv_bwd.py 2 20 $l; done; } > gpurun_out/conv_wgrad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/conv2_wgrad_3xtf32 -f \
    python tools/probes/one_conv_bwd.py 2 3 1 > gpurun_out/ncu_wgrad.log 2>&1
echo done
