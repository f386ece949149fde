"""Randomised cross-check of the oracle restatements against the reference's OWN classes, imported from /root/reference
(build container) or from its untouched copy oracle/_ref (oracle/make_ref.py; travels to the GPU box).  The committed
goldens of tests/golden/ are the portable pin, this is the wide one: several hundred random frames per kernel, bit for bit."""
import logging
import sys
import types

import numpy as np
import pandas as pd
import pytest

from oracle import bridge_oracle as B, golden_io, reference_loader as RL
from oracle.df_oracle import OracleCandleReconstructor, OracleIVInterpolator

pytestmark = pytest.mark.skipif(not RL.available(), reason="neither /root/reference nor oracle/_ref is present")


def _interp_frame(rng, n):
    step = rng.choice([60, 120, 3600, 7200])
    t0 = np.datetime64("2023-03-15T00:00:00", "us") + np.timedelta64(int(rng.integers(0, 3600)), "s")
    t = t0 + (np.arange(n) * step).astype("timedelta64[s]").astype("timedelta64[us]")
    keep = rng.random(n) >= rng.choice([0.0, 0.1, 0.3])
    keep[[0, -1]] = True
    t = t[keep]
    m = len(t)
    off = rng.random(m) < 0.03
    off[[0, -1]] = False
    t = t + np.where(off, 30_000_000, 0).astype("timedelta64[us]")
    def col(lo, hi, nan):
        v = rng.uniform(lo, hi, m)
        v[rng.random(m) < nan] = np.nan
        return v
    return pd.DataFrame({"symbol": "S", "date": t, "iv": col(0.2, 1.2, rng.choice([0.0, 0.05, 0.4])),
                         "underlying_price": col(2e4, 3e4, 0.02), "time_to_maturity": col(0.01, 1.0, 0.02),
                         "strike": 25000.0, "callput": "c", "interest_rate": 0.05, "mark_price": col(0.0, 0.2, 0.1),
                         "index_price": col(2e4, 3e4, 0.0), "volume": col(0, 10, 0.1), "quote_volume": col(0, 1e5, 0.1),
                         "record_time": [f"r{i}" for i in range(m)]})


def test_interp_oracle_equals_reference_on_random_frames():
    rng = np.random.default_rng(101)
    ora = OracleIVInterpolator()
    n_frames = 0
    for case in range(250):
        df = _interp_frame(rng, int(rng.integers(2, 200)))
        want = RL.reference_interpolate(df)
        got = ora.interpolate_symbol(df)
        try:
            golden_io.assert_frames_match(want, got)
        except AssertionError as e:
            raise AssertionError(f"case {case}: {e}") from None
        n_frames += want is not None
    assert n_frames > 150


def _candle_frame(rng, n):
    t0 = np.datetime64("2024-01-01T09:00:00", "us") + np.timedelta64(int(rng.integers(0, 600)), "s")
    gaps = rng.choice([60, 60, 60, 120, 300], n)
    t = t0 + np.cumsum(gaps).astype("timedelta64[s]").astype("timedelta64[us]")
    base = 100 + np.cumsum(rng.normal(0, 1, n))
    d = {"symbol": "S", "timestamp": t, "open": base, "high": base + np.abs(rng.normal(0, 1, n)),
         "low": base - np.abs(rng.normal(0, 1, n)), "close": base + rng.normal(0, 0.3, n), "volume": rng.exponential(50, n)}
    df = pd.DataFrame(d)
    for c in ("open", "high", "low", "close", "volume"):
        df.loc[rng.random(n) < rng.choice([0.0, 0.05, 0.3]), c] = np.nan
    return df


def test_candle_oracle_equals_reference_on_random_frames():
    rng = np.random.default_rng(202)
    n_frames = 0
    for case in range(250):
        df = _candle_frame(rng, int(rng.integers(1, 400)))
        freq = str(rng.choice(["5min", "15min", "7min", "1min", "3m"]))
        want = RL.reference_reconstruct(df, freq)
        got = OracleCandleReconstructor(freq).reconstruct_symbol_candles(df)
        try:
            golden_io.assert_frames_match(want, got)
        except AssertionError as e:
            raise AssertionError(f"case {case} ({freq}): {e}") from None
        n_frames += want is not None and len(want) > 0
    assert n_frames > 150


def test_bridge_oracle_equals_reference_on_random_frames():
    src = RL.REFERENCE_ROOT + "/src"
    if src not in sys.path:
        sys.path.insert(0, src)
    from data_bridge.ohlcv_converter import InterpolatedToOHLCVConverter
    logging.getLogger("data_bridge.ohlcv_converter").setLevel(logging.CRITICAL)
    rng = np.random.default_rng(303)
    for case in range(120):
        strat = str(rng.choice(["spread_simulation", "price_as_midpoint", "trend_following", "simple_spread"]))
        n = int(rng.integers(1, 400))
        price = float(rng.choice([25000.0, 0.37, 3.0, 64000.0]))
        base = price + np.cumsum(rng.normal(0, price * 0.002, n))
        vol = rng.exponential(5.0, n)
        if case % 3 == 1:
            # tidy prices (1 or 2 decimals) and 7-/8-decimal volumes: base * 1.0005 and the volumes land on decimal ties of the
            # rounded place, which is where Python's round (what the reference's object rows get) and numpy's part ways
            base = np.round(np.abs(base) + 0.5, int(rng.integers(1, 3)))
            vol = np.round(vol, int(rng.integers(7, 9)))
        base[rng.random(n) < 0.05] = np.nan
        base[rng.random(n) < 0.02] = 0.0
        vol[rng.random(n) < rng.choice([0.0, 0.2])] = np.nan
        df = pd.DataFrame({"symbol": "S", "timestamp": pd.date_range("2024-03-01", periods=n, freq="1min"), "underlying_price": base,
                           "mark_price": base, "index_price": base, "volume": vol})
        cfg = types.SimpleNamespace(data_bridge=types.SimpleNamespace(conversion_strategy=strat, spread_method="adaptive",
                                                                      enable_quality_checks=True,
                                                                      spread_parameters={"base_spread_percent": 0.002, "volatility_factor": 1.5}))
        seed = 5000 + case
        np.random.seed(seed)
        want = InterpolatedToOHLCVConverter(None, cfg)._generate_ohlcv_from_interpolated(df)
        rs = np.random.RandomState(seed)
        u, e, z = B.draw_tapes(rs, strat, base, vol)
        assert np.random.random_sample() == rs.random_sample(), (case, strat, "generator state")
        got = B.bridge_ohlcv(strat, base, vol, u, e, z)
        if want is None:
            assert len(got["row"]) == 0, (case, strat)
            continue
        for c in ("open", "high", "low", "close", "volume", "source_price"):
            assert np.array_equal(want[c].to_numpy(dtype=np.float64), got[c]), (case, strat, c)
