#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/bench_n_l1.json 2> gpurun_out/bench_n_l1.err
# launch list of one warm eager step + graph replays (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3200 --csv --log-file gpurun_out/launches_r01f.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --lanes 1 > gpurun_out/ncu_bench.log 2>&1
# full captures: conv.2 forward (kw-box kernel) in 3xTF32 and TF32, a cluster split-K GEMM, the short-sequence attention
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_kw -s 2 -c 1 -o gpurun_out/conv2_kw_3xtf32 -f \
    python tools/probes/one_conv.py 2 4 1 > gpurun_out/ncu_conv_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_kw -s 2 -c 1 -o gpurun_out/conv2_kw_tf32 -f \
    python tools/probes/one_conv.py 1 4 1 > gpurun_out/ncu_conv_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/gemm_264x512x512 -f \
    python tests/gpu_one_gemm.py 2 0 1 264 512 512 4 > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_small -s 4 -c 2 -o gpurun_out/attn_small -f \
    python tools/probes/attn_probe.py > gpurun_out/ncu_attn.log 2>&1
echo done
