"""The Arrow table API (iv_interpolation_b200/tables.py) on the CPU tier: grouping/sorting, column plumbing and table assembly are
the product's code; the launch is the oracle-backed stand-in of tests/fake_device.py.  Held to the reference's classes run symbol by
symbol on the same rows (where the reference is present) and to the DataFrame oracle otherwise."""
import numpy as np
import pandas as pd
import pyarrow as pa
import pytest
import torch

from iv_interpolation_b200 import columnar, tables, tapes
from oracle import reference_loader as RL
from oracle.df_oracle import OracleCandleReconstructor, OracleIVInterpolator
from tests import fake_device as FD

CPU = torch.device("cpu")


def _task1_table(seed=4, n=12, shuffle=True):
    tp = tapes.make_contract_tape(seed, n)
    frames = [tp.frame(i) for i in range(tp.n_sym)]
    frames[2].loc[3, "callput"] = None
    big = pd.concat(frames, ignore_index=True)
    if shuffle:
        big = big.iloc[np.random.default_rng(seed).permutation(len(big))].reset_index(drop=True)
    return frames, pa.Table.from_pandas(big, preserve_index=False)


def _want_interp(f):
    if RL.available():
        return RL.reference_interpolate(f)
    return OracleIVInterpolator().interpolate_symbol(f)


def test_interpolate_table_equals_per_symbol_reference_calls():
    frames, t = _task1_table()
    out, st = tables.interpolate_table(t, launch=FD.interp_linear_ffill, dev=CPU)
    assert out.column_names == ["date", "symbol", "iv", "underlying_price", "time_to_maturity", "strike", "callput", "interest_rate",
                                "mark_price", "index_price", "volume", "quote_volume", "record_time", "is_interpolated"]
    got = out.to_pandas()
    first_seen = list(dict.fromkeys(t["symbol"].to_pylist()))
    assert list(dict.fromkeys(got["symbol"])) == [s for s in first_seen if st[s]["status"] == "success"]
    for f in frames:
        sym = f["symbol"].iloc[0]
        want = _want_interp(f)
        g = got[got["symbol"] == sym]
        assert st[sym]["status"] == "success" and st[sym]["output_rows"] == len(want) == len(g) and st[sym]["input_rows"] == len(f)
        assert (g["date"].to_numpy() == want["date"].to_numpy()).all()
        for c in ("iv", "underlying_price", "time_to_maturity", "strike", "interest_rate", "mark_price", "index_price", "volume", "quote_volume"):
            assert np.array_equal(g[c].to_numpy(), want[c].to_numpy(dtype=np.float64), equal_nan=True), (sym, c)
        for c in ("callput", "record_time"):
            a, b = g[c].to_numpy(dtype=object), want[c].to_numpy(dtype=object)
            assert all((x == y) or (pd.isna(x) and pd.isna(y)) for x, y in zip(a, b)), (sym, c)
        assert not g["is_interpolated"].any()


def test_interpolate_table_guards_and_decimal_inputs():
    tp = tapes.make_contract_tape(9, 4)
    frames = [tp.frame(i) for i in range(4)]
    frames[1] = frames[1].iloc[:5]                                        # below min_points -> skipped (core.py:26-28)
    big = pd.concat(frames, ignore_index=True)
    t = pa.Table.from_pandas(big, preserve_index=False)
    t = t.set_column(t.column_names.index("strike"), "strike", t["strike"].cast(pa.decimal128(15, 4)))   # DECIMAL as stored
    out, st = tables.interpolate_table(t, launch=FD.interp_linear_ffill, dev=CPU)
    sym1 = frames[1]["symbol"].iloc[0]
    assert st[sym1]["status"] == "skipped" and st[sym1]["detail"] == "Insufficient data points"
    assert sum(s["status"] == "success" for s in st.values()) == 3 and out["strike"].type == pa.float64()
    with pytest.raises(KeyError):
        tables.interpolate_table(t.drop_columns(["iv"]), launch=FD.interp_linear_ffill, dev=CPU)


def test_reconstruct_table_equals_per_symbol_calls_and_db_precision():
    rng = np.random.default_rng(2)
    frames = []
    for k in range(6):
        n = int(rng.integers(3, 90))
        ts = pd.date_range("2024-03-01 09:02", periods=n, freq="1min")
        base = 100 + np.cumsum(rng.normal(0, 1, n))
        f = pd.DataFrame({"symbol": f"S{k}", "timestamp": ts, "open": base, "high": base + 1, "low": base - 1, "close": base + 0.25,
                          "volume": rng.exponential(5, n)})
        if k == 2:
            f.loc[4, "volume"] = np.nan
        frames.append(f)
    big = pd.concat(frames, ignore_index=True).iloc[rng.permutation(sum(len(f) for f in frames))].reset_index(drop=True)
    out, st = tables.reconstruct_table(pa.Table.from_pandas(big, preserve_index=False), "5min", launch=FD.ohlcv_reduce, dev=CPU)
    assert out.column_names == ["symbol", "timestamp", "open", "high", "low", "close", "volume", "frequency", "source_candles", "created_at"]
    got = out.to_pandas()
    for f in frames:
        sym = f["symbol"].iloc[0]
        want = RL.reference_reconstruct(f, "5min") if RL.available() else OracleCandleReconstructor("5min").reconstruct_symbol_candles(f)
        g = got[got["symbol"] == sym]
        if want is None or want.empty:
            assert len(g) == 0 and st[sym]["status"] != "success"
            continue
        assert st[sym] == {"symbol": sym, "status": "success", "input_candles": len(f), "output_candles": len(want)}
        assert (g["timestamp"].to_numpy() == want["timestamp"].to_numpy()).all()
        for c in ("open", "high", "low", "close", "volume"):
            assert np.array_equal(g[c].to_numpy(), want[c].to_numpy()), (sym, c)
    db = columnar.table_to_db_precision(out, columnar.CANDLE_DECIMALS)
    assert db["open"].type == pa.decimal128(15, 4) and db["volume"].type == pa.decimal128(20, 8) and db["symbol"].type == out["symbol"].type
    assert abs(float(db["close"][0].as_py()) - out["close"][0].as_py()) <= 5e-5
    with pytest.raises(ValueError):
        tables.reconstruct_table(pa.Table.from_pandas(big, preserve_index=False), "1h", launch=FD.ohlcv_reduce, dev=CPU)


def test_group_by_symbol_orders_and_drops_null_symbols():
    t = pa.table({"symbol": ["b", "a", None, "b", "a", "b"],
                  "date": pa.array(np.array([30, 10, 5, 20, None, 10], dtype=object), pa.timestamp("us"))})
    g = tables.group_by_symbol(t, "date")
    assert g.symbols.to_pylist() == ["b", "a"] and list(g.off) == [0, 3, 5] and g.unit == "us"
    assert list(g.order) == [5, 3, 0, 1, 4]                               # b: 10, 20, 30; a: 10, NaT last; the null symbol is gone
    s = pa.table({"symbol": ["x", "x", "y"], "date": pa.array([1, 2, 1], pa.timestamp("ns"))})
    assert tables.group_by_symbol(s, "date").order is None
