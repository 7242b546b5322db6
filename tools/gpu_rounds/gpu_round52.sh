#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_kw -s 2 -c 1 -o gpurun_out/conv4_kw_3xtf32 -f \
    python tools/probes/one_conv.py 2 4 3 > gpurun_out/ncu_conv4.log 2>&1
echo done
