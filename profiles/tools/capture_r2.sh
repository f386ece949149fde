#!/bin/bash
# Round-2 ncu passes (run only after the same bench command exited 0 without ncu):
#   capture_r2.sh <tag> launches   launch list of the default surface bench (per-launch gpu__time_duration)
#   capture_r2.sh <tag> full       one ncu --set full capture (with source) of the hot kernels at 600 stamps
set -e
TAG=${1:-r2}
WHAT=${2:-launches}
if [ "$WHAT" = launches ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py --skip-cpu --skip-e2e --skip-dropin --skip-stress --skip-contracts --steps 2 --warmup 1 > gpurun_out/ncu_launch_${TAG}.log 2>&1
  tail -c 300 gpurun_out/ncu_launch_${TAG}.log
else
  ncu --set full --clock-control none --import-source on \
      -k 'regex:^k_(bs_implied_vol|band_apply|spline_fit_eval|spline_strided|ohlc_time_major|interp_exec|ohlcv_reduce|bridge_ohlcv_seg)' -c 16 \
      -o gpurun_out/prof_${TAG} -f python bench.py --skip-cpu --skip-e2e --skip-dropin --skip-stress --skip-contracts --n-t 600 --steps 1 --warmup 0 > gpurun_out/ncu_full_${TAG}.log 2>&1
  ls -la gpurun_out/
fi
