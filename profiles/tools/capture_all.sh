#!/bin/bash
# Passes over one box (each only after the previous exited 0): the bench line, the CPU arm, the launch list of the same
# command (pass "bench"), and one ncu --set full capture of every kernel on the path (pass "ncu").  Outputs under
# gpurun_out/ (which merges back only below 64 MiB: the two halves are separate calls).
set -e
TAG=${1:-r1_h}
WHAT=${2:-bench}
if [ "$WHAT" = bench ]; then
  python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
  python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.json 2> gpurun_out/bench_${TAG}_reference_arm.err
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py --skip-cpu > gpurun_out/ncu_launch.log 2>&1
  tail -c 300 gpurun_out/bench_${TAG}.json
else
  ncu --set full --clock-control none --import-source on \
      -k 'regex:^k_(bs_implied_vol|spline_fit_eval|spline_strided|ohlc_time_major|interp_exec|interp_plan|ohlcv_reduce|bridge_ohlcv_seg|bridge_count_seg|bridge_check)' -c 24 \
      -o gpurun_out/prof_${TAG} -f python bench.py --skip-cpu --skip-e2e --n-t 600 --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out/
fi
