"""Generate tests/golden/*.npz by running the REFERENCE's own classes.  TEST INFRASTRUCTURE ONLY.

Run in the build container, where /root/reference is mounted:

    python -m oracle.make_golden

The reference modules are imported from /root/reference (never copied); Task 1 runs under the
``fillna(method=)`` shim explained in oracle/reference_loader.py.  Inputs AND outputs are
stored, so the fixtures do not depend on RNG stream stability.  Versions used are recorded in
tests/golden/MANIFEST.json.  The known-answer cases are the ones SURVEY.md §8(c) lists
(A-1..A-8, B-1..B-7) plus seeded tapes and edge cases (unsorted input, duplicates, NaT, null
symbols, non-finite values, pre-epoch stamps, other datetime units).
"""
from __future__ import annotations

import json
import logging
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from iv_interpolation_b200 import tapes  # noqa: E402
from oracle import golden_io, reference_loader as R  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _task1_frame(dates, iv, S=None, ttm=None, **over):
    n = len(dates)
    d = {
        "symbol": "btc-20mar23-24500-c",
        "date": pd.to_datetime(list(dates), format="mixed") if isinstance(dates[0], str) else pd.to_datetime(dates),
        "iv": np.asarray(iv, dtype=float),
        "underlying_price": np.asarray(S if S is not None else 100.0 + np.arange(n), dtype=float),
        "time_to_maturity": np.asarray(ttm if ttm is not None else 1.0 - 0.05 * np.arange(n), dtype=float),
        "strike": 24500.0, "callput": "c", "interest_rate": 0.01,
        "mark_price": 0.1 + 0.01 * np.arange(n), "index_price": 100.0 + np.arange(n),
        "volume": np.arange(n, dtype=float), "quote_volume": 2.0 * np.arange(n),
        "record_time": [f"r{i}" for i in range(n)],
    }
    d.update(over)
    return pd.DataFrame(d)


def interp_cases():
    """(name, frame, kwargs) for IVInterpolator(**kwargs).interpolate_symbol(frame)."""
    day = "2023-03-15 "
    C = []
    kw2 = {"min_points": 2}
    C.append(("A1_hourly3", _task1_frame([day + "00:00", day + "01:00", day + "02:00"], [.5, .8, .2],
                                          S=[100, 101, 102], ttm=[1.0, .95, .9]), kw2))
    C.append(("A2_offgrid_row_dropped", _task1_frame([day + "00:00", day + "00:30:30", day + "01:00"], [.5, 9.9, .8]), kw2))
    C.append(("A3_grid_anchored_at_min_seconds", _task1_frame([day + "00:00:17", day + "00:02:17", day + "00:03:00"], [.5, .6, .7]), kw2))
    C.append(("A3b_anchor", _task1_frame([day + "00:00:17", day + "00:01:17", day + "00:02:17", day + "00:03:00"], [.5, .6, .7, .9]), kw2))
    C.append(("A4_last_offgrid_flat_tail", _task1_frame([day + "00:00", day + "00:02", day + "00:04:30"], [.5, .7, 9.9]), kw2))
    C.append(("A5_duplicate_stamp", _task1_frame([day + "00:00", day + "00:02", day + "00:02", day + "00:04"], [.5, .7, .9, 1.1]), kw2))
    C.append(("A6_nan_lead_mid_tail", _task1_frame([day + f"00:{2*i:02d}" for i in range(5)], [np.nan, .4, np.nan, .8, np.nan]), kw2))
    t31 = pd.Timestamp("2023-03-15")
    C.append(("A7_span_31d_none", _task1_frame([t31, t31 + pd.Timedelta(days=31)], [.5, .6]), kw2))
    C.append(("A7_min_points_default_2rows_none", _task1_frame([day + "00:00", day + "01:00"], [.5, .6]), {}))
    C.append(("A7_method_time_none", _task1_frame([day + "00:00", day + "01:00", day + "02:00"], [.5, .8, .2]), {"min_points": 2, "method": "time"}))
    C.append(("A7_method_index", _task1_frame([day + "00:00", day + "01:00", day + "02:00"], [.5, .8, .2]), {"min_points": 2, "method": "index"}))
    C.append(("A7_method_values", _task1_frame([day + "00:00", day + "01:00", day + "02:00"], [.5, .8, .2]), {"min_points": 2, "method": "values"}))
    rng = np.random.default_rng(8)
    C.append(("A8_25_hourly_random", _task1_frame(pd.date_range("2023-03-15", periods=25, freq="1h"), rng.uniform(.2, 1.2, 25),
                                                   S=25000 + np.cumsum(rng.normal(0, 50, 25)), ttm=np.linspace(.1, .09, 25)), {}))
    # unsorted, unique stamps
    f = _task1_frame(pd.date_range("2023-03-15", periods=12, freq="1h"), rng.uniform(.2, 1.2, 12))
    C.append(("unsorted_rows", f.iloc[rng.permutation(12)].reset_index(drop=True), {}))
    # column-wise ffill: NaN inside fill columns, null strings, null leading symbol
    f = _task1_frame(pd.date_range("2023-03-15", periods=12, freq="30min"), rng.uniform(.2, 1.2, 12))
    f.loc[[0, 4], "volume"] = np.nan
    f.loc[[3], "strike"] = np.nan
    f["callput"] = f["callput"].astype(object)
    f.loc[[2, 3], "callput"] = None
    f["record_time"] = f["record_time"].astype(object)
    f.loc[[0], "record_time"] = None
    C.append(("columnwise_ffill_nulls", f, {}))
    g = f.copy()
    g["symbol"] = g["symbol"].astype(object)
    g.loc[[0, 1], "symbol"] = None
    C.append(("leading_null_symbol_rows_dropped", g, {}))
    # NaT date
    g = _task1_frame(pd.date_range("2023-03-15", periods=11, freq="1h"), rng.uniform(.2, 1.2, 11))
    g.loc[4, "date"] = pd.NaT
    C.append(("nat_date_row", g, {}))
    # non-finite values
    g = _task1_frame(pd.date_range("2023-03-15", periods=11, freq="10min"), rng.uniform(.2, 1.2, 11))
    g.loc[2, "iv"] = np.inf
    g.loc[5, "underlying_price"] = -np.inf
    g.loc[6, "underlying_price"] = np.nan
    C.append(("inf_values", g, {}))
    g = _task1_frame(pd.date_range("2023-03-15", periods=11, freq="10min"), rng.uniform(.2, 1.2, 11))
    g.loc[2, "iv"] = np.inf
    g.loc[3, "iv"] = np.nan
    g.loc[4, "iv"] = -np.inf
    C.append(("inf_to_minus_inf_interior_nan", g, {}))
    # units / column subsets / extras
    g = _task1_frame(pd.date_range("2023-03-15", periods=10, freq="1h"), rng.uniform(.2, 1.2, 10))
    g["date"] = g["date"].astype("datetime64[ns]")
    C.append(("unit_ns", g, {}))
    g = g.copy()
    g["date"] = g["date"].astype("datetime64[s]")
    C.append(("unit_s", g, {}))
    g = _task1_frame(pd.date_range("2023-03-15", periods=10, freq="1h"), rng.uniform(.2, 1.2, 10))
    C.append(("missing_optional_cols", g.drop(columns=["mark_price", "record_time"]), {}))
    g2 = g.copy()
    g2["extra"] = np.arange(10.0)
    C.append(("extra_column_not_filled", g2, {}))
    C.append(("missing_symbol_none", g.drop(columns=["symbol"]), {}))
    C.append(("missing_iv_none", g.drop(columns=["iv"]), {}))
    g3 = g.copy()
    g3["date"] = g3["date"].astype(str)
    C.append(("string_dates", g3, {}))
    g4 = g.copy()
    g4["iv"] = np.nan
    C.append(("all_nan_iv_none", g4, {}))
    # exactly 30 days passes (43 201 rows); 100 000-point cap cannot be reached inside 30 d at 1 min
    C.append(("span_exactly_30d", _task1_frame([t31, t31 + pd.Timedelta(days=15), t31 + pd.Timedelta(days=30)], [.5, .9, .6]), kw2))
    # seeded tapes: config 1 (hourly, 25 rows) and a sparse config-5-like one
    tp = tapes.make_contract_tape(0, 16)
    for i in range(tp.n_sym):
        C.append((f"tape_cfg1_seed0_{i}", tp.frame(i), {}))
    tp = tapes.make_contract_tape(4, 8, rows_per_contract=73, spacing_us=20 * tapes.MINUTE_US, nan_frac=0.3, missing_frac=0.1)
    for i in range(tp.n_sym):
        C.append((f"tape_cfg5_seed4_{i}", tp.frame(i), {"min_points": 5}))
    return C


def _task2_frame(ts, n=None, **over):
    ts = pd.to_datetime(ts)
    n = len(ts)
    i = np.arange(n, dtype=float)
    d = {"symbol": "btc-20mar23-24500-c", "timestamp": ts, "open": 100 + i, "high": 101 + i,
         "low": 99 + i, "close": 100.5 + i, "volume": np.ones(n)}
    d.update(over)
    return pd.DataFrame(d)


def candle_cases():
    """(name, frame, target_frequency)."""
    C = []
    rng = np.random.default_rng(9)
    C.append(("B1_23rows_from_0902", _task2_frame(pd.date_range("2023-03-15 09:02", periods=23, freq="1min")), "5min"))
    f = _task2_frame(pd.date_range("2023-03-15 09:00", periods=10, freq="1min"))
    f.loc[0, "open"] = np.nan; f.loc[4, "close"] = np.nan; f.loc[4, "high"] = np.nan; f.loc[2, "volume"] = np.nan
    C.append(("B2_nan_skipping", f, "5min"))
    ts = list(pd.date_range("2023-03-15 09:00", periods=4, freq="1min")) \
        + list(pd.date_range("2023-03-15 09:05", periods=4, freq="1min")) + [pd.Timestamp("2023-03-15 09:08:30")] \
        + list(pd.date_range("2023-03-15 09:10", periods=5, freq="1min")) + [pd.Timestamp("2023-03-15 09:14:30")]
    C.append(("B3_counts_offminute_rows", _task2_frame(ts), "5min"))
    C.append(("B4_7min_epoch_anchor", _task2_frame(pd.date_range("2023-03-15 00:00", periods=30, freq="1min")), "7min"))
    f = _task2_frame(pd.date_range("2023-03-15", periods=2000 * 5, freq="1min"), volume=rng.exponential(50.0, 10000) * 10.0 ** rng.integers(-6, 7, 10000))
    C.append(("B5_kahan_2000_buckets", f, "5min"))
    f = _task2_frame(pd.date_range("2023-03-15 09:00", periods=10, freq="1min"))
    f.loc[0:4, "open"] = np.nan; f.loc[0:4, "volume"] = np.nan
    C.append(("B6_all_nan_open_volume", f, "5min"))
    C.append(("freq_15m", _task2_frame(pd.date_range("2023-03-15 09:00", periods=50, freq="1min")), "15m"))
    C.append(("freq_1min_identity", _task2_frame(pd.date_range("2023-03-15 09:00", periods=7, freq="1min")), "1min"))
    C.append(("too_few_rows_none", _task2_frame(pd.date_range("2023-03-15 09:00", periods=3, freq="1min")), "5min"))
    C.append(("no_complete_bucket_empty_frame", _task2_frame(pd.date_range("2023-03-15 09:03", periods=6, freq="1min")), "5min"))
    C.append(("missing_volume_none", _task2_frame(pd.date_range("2023-03-15 09:00", periods=10, freq="1min")).drop(columns=["volume"]), "5min"))
    C.append(("missing_symbol_none", _task2_frame(pd.date_range("2023-03-15 09:00", periods=10, freq="1min")).drop(columns=["symbol"]), "5min"))
    C.append(("pre_epoch_7min", _task2_frame(pd.date_range("1969-12-31 23:40", periods=40, freq="1min")), "7min"))
    f = _task2_frame(pd.date_range("2023-03-15 09:00", periods=20, freq="1min"))
    f = f.iloc[rng.permutation(20)].reset_index(drop=True)
    C.append(("unsorted_rows", f, "5min"))
    f = _task2_frame(pd.date_range("2023-03-15 09:00", periods=12, freq="1min"))
    f.loc[3, "timestamp"] = pd.NaT
    C.append(("nat_row", f, "5min"))
    f = _task2_frame(pd.date_range("2023-03-15 09:00", periods=10, freq="1min"))
    f["symbol"] = f["symbol"].astype(object)
    f.loc[0, "symbol"] = None; f.loc[5:9, "symbol"] = None
    C.append(("null_symbols", f, "5min"))
    f = _task2_frame(pd.date_range("2023-03-15 09:00", periods=15, freq="1min"))
    f.loc[1, "high"] = np.inf; f.loc[2, "low"] = -np.inf; f.loc[3, "volume"] = np.inf; f.loc[4, "volume"] = -np.inf
    f.loc[7, "volume"] = np.inf; f.loc[8, "volume"] = 3.0; f.loc[11, "volume"] = -np.inf
    C.append(("inf_values", f, "5min"))
    f = _task2_frame(pd.date_range("2023-03-15 09:00", periods=15, freq="1min"))
    f["timestamp"] = f["timestamp"].astype("datetime64[ns]")
    C.append(("unit_ns", f, "5min"))
    f2 = f.copy(); f2["timestamp"] = f2["timestamp"].astype("datetime64[s]")
    C.append(("unit_s", f2, "5min"))
    f = _task2_frame(pd.date_range("2023-03-15 09:00:00", periods=200, freq="1s"))
    C.append(("seconds_to_1min", f, "1min"))
    ct = tapes.make_candle_tape(1, 6, rows_per_symbol=1441, gap_frac=0.02, nan_frac=0.02)
    for i in range(ct.n_sym):
        for fr in ("5min", "7min", "1min", "60min"):
            C.append((f"tape_seed1_{i}_{fr}", ct.frame(i), fr))
    return C


def greeks_cases():
    return [dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2, option_type="call"),
            dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2, option_type="put"),
            dict(S=4500.0, K=5200.0, T=0.25, r=0.05, sigma=0.35, option_type="call"),
            dict(S=4500.0, K=3900.0, T=1.5, r=0.03, sigma=0.8, option_type="put")]


def main():
    assert R.available(), "run in the build container where /root/reference is mounted"
    logging.disable(logging.CRITICAL)
    os.makedirs(GOLDEN, exist_ok=True)
    manifest = {"pandas": pd.__version__, "numpy": np.__version__, "python": sys.version.split()[0],
                "reference_root": R.REFERENCE_ROOT, "task1_shim": "Series.fillna(method='ffill') -> .ffill()"}

    d, names = {}, []
    for name, frame, kw in interp_cases():
        out = R.reference_interpolate(frame, **kw)
        golden_io.frame_to_dict(f"{name}/in", frame, d)
        golden_io.frame_to_dict(f"{name}/out", out, d)
        d[f"{name}/kwargs"] = np.array(json.dumps(kw))
        names.append(name)
    d["names"] = np.array(json.dumps(names))
    np.savez_compressed(os.path.join(GOLDEN, "interp_cases.npz"), **d)
    manifest["interp_cases"] = len(names)

    d, names = {}, []
    for name, frame, freq in candle_cases():
        out = R.reference_reconstruct(frame, freq)
        if out is not None:
            out = out.drop(columns=["created_at"])        # wall-clock, not part of parity
        golden_io.frame_to_dict(f"{name}/in", frame, d)
        golden_io.frame_to_dict(f"{name}/out", out, d)
        d[f"{name}/freq"] = np.array(freq)
        cls = R.load_reconstructor_cls()
        rec = cls(freq)
        d[f"{name}/valid_in"] = np.array(bool(rec.validate_candle_data(frame)) if all(c in frame.columns for c in ("open", "high", "low", "close", "volume")) else False)
        d[f"{name}/valid_out"] = np.array(bool(rec.validate_candle_data(out)) if out is not None else False)
        names.append(name)
    d["names"] = np.array(json.dumps(names))
    np.savez_compressed(os.path.join(GOLDEN, "candle_cases.npz"), **d)
    manifest["candle_cases"] = len(names)

    # B-7: frequency strings the reference refuses (candle core.py:27-34)
    cls = R.load_reconstructor_cls()
    freq_kat = {}
    for fr in ("5min", "15m", "1min", "60min", "0min", "1h", "60s", "5", "", "min"):
        try:
            freq_kat[fr] = cls(fr).frequency_minutes
        except ValueError:
            freq_kat[fr] = "ValueError"
    G = R.load_greeks_cls()
    gk = []
    for c in greeks_cases():
        res = G.calculate_greeks(**c)
        gk.append({"in": c, "out": {k: float(v) for k, v in res.items()}})
    with open(os.path.join(GOLDEN, "kats.json"), "w") as f:
        json.dump({"parse_frequency": freq_kat, "greeks": gk}, f, indent=1)
    with open(os.path.join(GOLDEN, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    for fn in sorted(os.listdir(GOLDEN)):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)))


if __name__ == "__main__":
    main()
