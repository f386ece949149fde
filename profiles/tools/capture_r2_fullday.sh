#!/bin/bash
# ncu --set full of the solver and the banded strike spline at the FULL day (23 400 stamps, 1.404 G quotes), one launch each
set -e
TAG=${1:-r2c}
ncu --set full --clock-control none --import-source on -k 'regex:^k_(bs_implied_vol_ladder2|band_apply_mma)$' -c 2 \
    -o gpurun_out/prof_${TAG}_fullday -f python bench.py --skip-cpu --skip-e2e --skip-dropin --skip-stress --skip-contracts --steps 1 --warmup 0 > gpurun_out/ncu_fullday_${TAG}.log 2>&1
ls -la gpurun_out/prof_${TAG}_fullday.ncu-rep
