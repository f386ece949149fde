#!/usr/bin/env python
"""EXPERIMENT: does the banded strike spline care how far apart the slices of one ladder lie in memory?
Same ladder, same number of slices; `stride` = distance (in slices) between consecutive slices of the batch order."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from iv_interpolation_b200 import device as D

dev = torch.device("cuda:0")
n_strike, n_q, n_sl = 2000, 512, 240000
rng = np.random.default_rng(0)
lad = 2000.0 + np.cumsum(rng.uniform(0.9, 1.1, n_strike))
xq = np.linspace(lad[0] * 1.001, lad[-1] * 0.999, n_q)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
y = torch.rand(n_sl * n_strike, dtype=torch.float64, device=dev) * 0.1 + 0.2
knot_off = torch.arange(n_sl + 1, dtype=torch.int64, device=dev) * n_strike
ladder_off = t(np.array([0, n_strike], np.int64))
x_off = torch.zeros(n_sl, dtype=torch.int64, device=dev)
slice_ladder = torch.zeros(n_sl, dtype=torch.int32, device=dev)
lad_slice_off = t(np.array([0, n_sl], np.int64))
lu = D.spline_ladder_factor(ladder_off, t(lad), max_knots=n_strike)
band = D.spline_band_build(ladder_off, t(lad), t(xq), extrap=1)
out = torch.empty(n_sl * n_q, dtype=torch.float64, device=dev)
fb = (torch.zeros(n_sl, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
for stride in (1, 30):
    order = np.arange(n_sl).reshape(-1, stride).T.reshape(-1).astype(np.int32) if stride > 1 else np.arange(n_sl, dtype=np.int32)
    order = t(order)
    for _ in range(2):
        D.spline_fit_eval_banded(knot_off, x_off, slice_ladder, order, lad_slice_off, n_sl, ladder_off, t(lad), lu, y, t(xq), band, max_knots=n_strike, out=out, fb=fb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        D.spline_fit_eval_banded(knot_off, x_off, slice_ladder, order, lad_slice_off, n_sl, ladder_off, t(lad), lu, y, t(xq), band, max_knots=n_strike, out=out, fb=fb)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"stride {stride}: {ms:.3f} ms per {n_sl} slices -> {ms * 702000 / n_sl:.2f} ms per day, flagged {int(fb[1].item())}")
