#!/bin/bash
# One parameterised GPU-box job: `gpurun -- 'bash tools/gpu_job.sh <step> [<step> ...]'`.  Every step writes into
# gpurun_out/ (scratch); what is worth keeping is copied to profiles/ by hand afterwards.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
PY=python
for step in "$@"; do
  echo "=== $step ($(date +%T))"
  case "$step" in
    pytest)      timeout 1500 $PY -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log ;;
    pytest_new)  timeout 1200 $PY -m pytest tests -m gpu -q -k "full_tensor or decision_flips or unchanged_reference or two_gpu or empty_transcript or varying_shapes or checkpoint" > $OUT/pytest_new.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_new.log; tail -15 $OUT/pytest_new.log ;;
    precision)   timeout 900 $PY tools/probes/precision_table.py > $OUT/precision_table.log 2>&1; tail -80 $OUT/precision_table.log ;;
    bench)       timeout 900 $PY bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json; tail -3 $OUT/bench.err ;;
    bench_tf32)  MTL_GEMM_MODE=1 timeout 600 $PY bench.py --no-cpu-baseline > $OUT/bench_tf32.json 2> $OUT/bench_tf32.err; tail -c 1500 $OUT/bench_tf32.json ;;
    bench_1lane) timeout 600 $PY bench.py --no-cpu-baseline --lanes 1 > $OUT/bench_1lane.json 2> $OUT/bench_1lane.err; tail -c 1500 $OUT/bench_1lane.json ;;
    bench_ref)   timeout 900 $PY bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 1500 $OUT/bench_ref.json ;;
    timeline)    timeout 600 $PY bench.py --no-cpu-baseline --steps 4 --warmup 3 --timeline $OUT/timeline.csv.gz > $OUT/bench_tl.json 2> $OUT/bench_tl.err; $PY tools/launch_summary.py $OUT/timeline.csv.gz > $OUT/timeline_summary.txt 2>&1; head -40 $OUT/timeline_summary.txt ;;
    smoke)       timeout 600 $PY -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log ;;
    memcheck)    timeout 1500 compute-sanitizer --tool memcheck --log-file $OUT/sanitizer_memcheck_smoke.log $PY -c 'import __graft_entry__ as g; g.smoke()' > $OUT/sanitizer_memcheck_smoke.out 2>&1; tail -5 $OUT/sanitizer_memcheck_smoke.log; tail -2 $OUT/sanitizer_memcheck_smoke.out ;;
    racecheck)   timeout 1500 compute-sanitizer --tool racecheck --log-file $OUT/sanitizer_racecheck_smoke.log $PY -c 'import __graft_entry__ as g; g.smoke()' > $OUT/sanitizer_racecheck_smoke.out 2>&1; tail -5 $OUT/sanitizer_racecheck_smoke.log; tail -2 $OUT/sanitizer_racecheck_smoke.out ;;
    synccheck)   timeout 1500 compute-sanitizer --tool synccheck --log-file $OUT/sanitizer_synccheck_smoke.log $PY -c 'import __graft_entry__ as g; g.smoke()' > $OUT/sanitizer_synccheck_smoke.out 2>&1; tail -5 $OUT/sanitizer_synccheck_smoke.log ;;
    memcheck_ops) timeout 1800 compute-sanitizer --tool memcheck --log-file $OUT/sanitizer_memcheck_ops.log $PY -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "gemm or conv3x3 or attention" > $OUT/sanitizer_memcheck_ops.out 2>&1; tail -5 $OUT/sanitizer_memcheck_ops.log; tail -3 $OUT/sanitizer_memcheck_ops.out ;;
    racecheck_ops) timeout 1800 compute-sanitizer --tool racecheck --log-file $OUT/sanitizer_racecheck_ops.log $PY -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "gemm or conv3x3 or attention" > $OUT/sanitizer_racecheck_ops.out 2>&1; tail -5 $OUT/sanitizer_racecheck_ops.log; tail -3 $OUT/sanitizer_racecheck_ops.out ;;
    eager)       timeout 900 $PY tools/gpu_eager_baseline.py > $OUT/gpu_eager.json 2> $OUT/gpu_eager.err; tail -c 1500 $OUT/gpu_eager.json; tail -3 $OUT/gpu_eager.err ;;
    *)           echo "running custom: $step"; timeout 1500 bash -c "$step" ;;
  esac
done
echo "=== done ($(date +%T))"
