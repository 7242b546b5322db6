#!/bin/bash
set -u
mkdir -p gpurun_out
{ MTL_CONV_KW_SEL=1 timeout 120 python tools/probes/ws_dump.py a
  MTL_CONV_KW_SEL=0 timeout 120 python tools/probes/ws_dump.py b
  timeout 300 python tools/probes/ws_diff.py
} > gpurun_out/ws_diff.log 2>&1
echo done
