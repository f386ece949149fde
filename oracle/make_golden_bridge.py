"""Golden vectors for the data bridge, produced by the reference's OWN code.  TEST INFRASTRUCTURE ONLY.

Runs in the build container only (needs /root/reference):  ``python -m oracle.make_golden_bridge``
  * imports ``data_bridge.ohlcv_converter.InterpolatedToOHLCVConverter`` from /root/reference/src unmodified and
    calls ``_generate_ohlcv_from_interpolated`` with numpy's global generator seeded per case;
  * ``complete_pipeline.py`` cannot be imported here (psycopg2, sqlalchemy ... are absent), so its
    ``_generate_ohlcv_candles`` is compiled FROM THE FILE WHERE IT LIES (ast extraction at run time, nothing copied
    into this repository) and called the same way;
  * records inputs, the outputs, and the (u, e, z) tapes ``oracle.bridge_oracle.draw_tapes`` draws from an identically
    seeded ``RandomState`` -> ``tests/golden/bridge_cases.npz``.
``tests/test_oracle_golden.py`` then holds ``oracle.bridge_oracle.bridge_ohlcv`` to these bit for bit.
"""
from __future__ import annotations

import ast
import json
import os
import sys
import types

import numpy as np
import pandas as pd

from oracle import bridge_oracle as B
from oracle.reference_loader import REFERENCE_ROOT

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bridge_cases.npz")


def _converter(strategy: str, params=None):
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    from data_bridge.ohlcv_converter import InterpolatedToOHLCVConverter

    cfg = types.SimpleNamespace(data_bridge=types.SimpleNamespace(
        conversion_strategy=strategy, spread_method="adaptive", enable_quality_checks=True,
        spread_parameters=params if params is not None else {"base_spread_percent": 0.002, "volatility_factor": 1.5}))
    return InterpolatedToOHLCVConverter(None, cfg)


def _pipeline_function():
    path = os.path.join(REFERENCE_ROOT, "complete_pipeline.py")
    tree = ast.parse(open(path).read(), path)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "_generate_ohlcv_candles":
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"pd": pd, "np": np}
            exec(compile(mod, path, "exec"), ns)
            return ns["_generate_ohlcv_candles"]
    raise RuntimeError("_generate_ohlcv_candles not found")


def make_frame(seed: int, n: int, nan_base: float, bad_volume: float, price: float = 25000.0):
    rng = np.random.default_rng(seed)
    ts = pd.date_range("2024-03-01", periods=n, freq="1min")
    base = price + np.cumsum(rng.normal(0, price * 0.002, n))
    volume = rng.exponential(5.0, n)
    base[rng.random(n) < nan_base] = np.nan
    base[rng.random(n) < nan_base / 2] = -1.0
    base[rng.random(n) < nan_base / 2] = 0.0
    volume[rng.random(n) < bad_volume] = np.nan
    volume[rng.random(n) < bad_volume / 2] = 0.0
    volume[rng.random(n) < bad_volume / 2] = -2.0
    return pd.DataFrame({"symbol": "SYM", "timestamp": ts, "underlying_price": base, "mark_price": base + 1.0,
                         "index_price": base - 1.0, "volume": volume, "is_interpolated": False, "batch_id": 1})


def main():
    cases = []
    out = {}
    pipe = _pipeline_function()
    specs = []
    for strat in ("spread_simulation", "price_as_midpoint", "trend_following", "simple_spread", "price_midpoint", "pipeline"):
        specs += [(strat, 7, 40, 0.0, 0.0, 25000.0), (strat, 8, 300, 0.05, 0.10, 25000.0), (strat, 9, 500, 0.10, 0.0, 0.37),
                  (strat, 10, 1500, 0.02, 0.30, 64000.0), (strat, 11, 3, 0.0, 1.0, 100.0), (strat, 12, 64, 0.9, 0.5, 2.5)]
    specs.append(("spread_simulation", 13, 400, 0.05, 0.05, 1e-3))      # prices where the 4-decimal rounding bites
    specs.append(("trend_following", 14, 400, 0.05, 0.05, 1e-3))
    # tidy prices (1 and 2 decimals) and 7-/8-decimal volumes: base +- 0.05 % and the volumes sit on decimal ties of the rounded
    # place, where Python's round (the reference's rows are object Series: Python floats) and numpy's differ
    for j, strat in enumerate(("simple_spread", "price_as_midpoint", "spread_simulation", "trend_following", "pipeline")):
        specs.append((strat, 20 + j, 300, 0.03, 0.10, -(1 + j % 2)))
    for k, (strat, seed, n, nan_base, bad_vol, price) in enumerate(specs):
        if price < 0:
            df = make_frame(seed, n, nan_base, bad_vol, 100.0)
            dec = int(-price)
            tidy = np.round(np.abs(df["underlying_price"].to_numpy()) + 0.5, dec)
            keep = df["underlying_price"].to_numpy()
            df["underlying_price"] = np.where(np.isnan(keep) | (keep <= 0), keep, tidy)
            df["mark_price"], df["index_price"] = df["underlying_price"] + 1.0, df["underlying_price"] - 1.0
            df["volume"] = np.where(df["volume"].to_numpy() > 0, np.round(df["volume"].to_numpy(), 7 + seed % 2), df["volume"].to_numpy())
        else:
            df = make_frame(seed, n, nan_base, bad_vol, price)
        np.random.seed(1000 + seed)
        if strat == "pipeline":
            res = pipe(None, df.rename(columns={"timestamp": "date"}))
            if res.empty:
                res = None
        else:
            res = _converter(strat)._generate_ohlcv_from_interpolated(df)
        rs = np.random.RandomState(1000 + seed)
        base_col = df["underlying_price"].to_numpy()
        if strat == "pipeline":        # complete_pipeline.py:478: underlying_price or mark_price or index_price (NaN is truthy, 0.0 is not)
            base_col = B.pipeline_base(base_col, df["mark_price"].to_numpy(), df["index_price"].to_numpy())
        u, e, z = B.draw_tapes(rs, strat, base_col, df["volume"].to_numpy())
        # the oracle's tapes must leave the generator where the reference left the global one
        assert np.random.random_sample() == rs.random_sample(), (strat, seed, "generator state diverged")
        key = f"case{k}"
        out[key + "/base"] = base_col
        out[key + "/volume"] = df["volume"].to_numpy()
        out[key + "/ts"] = df["timestamp"].to_numpy().astype("datetime64[us]").view(np.int64)
        out[key + "/u"], out[key + "/e"], out[key + "/z"] = u, e, z
        if res is None:
            out[key + "/none"] = np.array(1)
        else:
            tcol = "date" if strat == "pipeline" else "timestamp"
            for c in ("open", "high", "low", "close", "volume", "source_price"):
                out[key + "/out_" + c] = res[c].to_numpy(dtype=np.float64)
            out[key + "/out_ts"] = pd.to_datetime(res[tcol if tcol in res.columns else "timestamp"]).to_numpy().astype("datetime64[us]").view(np.int64)
            if strat != "pipeline":
                out[key + "/quality"] = np.array(json.dumps(_converter(strat)._validate_ohlcv_quality(res)))
        cases.append({"key": key, "strategy": strat, "seed": seed, "n": n, "rows_out": 0 if res is None else len(res)})
    out["manifest"] = np.array(json.dumps({"cases": cases, "numpy": np.__version__, "pandas": pd.__version__,
                                           "reference": "src/data_bridge/ohlcv_converter.py:138-412, complete_pipeline.py:473-510"}))
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(cases)} cases, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
