"""Arrow tables in, Arrow tables out: the hot path without a DataFrame (and without a Python loop) per symbol.

The reference reaches its classes through PostgreSQL: ``SELECT ... WHERE symbol = %s ORDER BY date`` per contract
(reference src/interpolation/batch_processor.py:76-87, src/candle_reconstruction/core.py:215-223), then ``iterrows`` +
``execute_values`` on the way out (:167-187, :284-320) — SURVEY.md section 8 f-1.  This module is the columnar
counterpart: ONE table holding every contract's rows (any row order; a ``symbol`` column says whose they are) goes to the
device as structure-of-arrays columns, one launch per stage runs all contracts, and the results come back as one table
with the reference's column names and order.  Strings never reach the device: ``symbol`` travels as dictionary codes,
the other text columns are gathered on the host by Arrow's ``take`` from the source-row index kernel A returns.

    interpolate_table   Task 1   IVInterpolator.interpolate_symbol for every symbol        (core.py:16-85)
    bridge_table        bridge   InterpolatedToOHLCVConverter._generate_ohlcv_from_...     (ohlcv_converter.py:138-263)
    reconstruct_table   Task 2   CandleReconstructor.reconstruct_symbol_candles            (candle core.py:36-106)
    pipeline_table      all three chained on the device (complete_pipeline.py:740-831), only tables asked for leave it

``iv_interpolation_b200/__main__.py`` is the DB-free command line over these (Parquet in, Parquet out);
``columnar.py`` holds the DECIMAL precisions of the reference's DDL for the files it writes.
"""
from __future__ import annotations

from dataclasses import dataclass
from datetime import datetime
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

from . import frames as F
from ._lib import SYM_OK, IvxError
from .hostpath import I64_MAX, _to_device, d2h_many

TASK1_ORDER = ("date", "symbol", "iv", "underlying_price", "time_to_maturity", "strike", "callput", "interest_rate",
               "mark_price", "index_price", "volume", "quote_volume", "record_time")            # core.py:54-55 merge order
_STATUS_REASON = {1: "Insufficient data points", 2: "Time span exceeds 30 days", 3: "Too many interpolation points",
                  4: "No valid data after interpolation", 5: "No valid dates"}


@dataclass
class SymbolGroups:
    symbols: pa.Array            # one entry per group (dictionary of the symbol column, first-appearance order)
    off: np.ndarray              # int64[n_sym+1] into the ordered rows
    order: Optional[np.ndarray]  # row permutation that groups by symbol and sorts by time (None: already so)
    ticks: np.ndarray            # int64 stamps, ordered
    unit: str
    tz: Optional[str]


def _f64(col: pa.ChunkedArray) -> np.ndarray:
    a = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    if not pa.types.is_float64(a.type):
        a = a.cast(pa.float64())                        # DECIMAL(p, s) / integers -> float8, as pd.read_sql coerces them
    if a.null_count:
        a = a.fill_null(float("nan"))
    return np.ascontiguousarray(a.to_numpy(zero_copy_only=False))


def group_by_symbol(table: pa.Table, time_column: str) -> SymbolGroups:
    """Rows grouped by ``symbol`` (first appearance order) and sorted by ``time_column`` inside a group, NaT last — the
    table-wide equivalent of the reference's per-symbol ``ORDER BY`` / ``sort_values``.  Rows with a null symbol are dropped."""
    sym = table[time_column]  # noqa: F841  (KeyError early when the stamp column is missing)
    sym = table["symbol"].combine_chunks()
    enc = sym if pa.types.is_dictionary(sym.type) else pc.dictionary_encode(sym)
    codes = enc.indices.fill_null(-1).to_numpy(zero_copy_only=False).astype(np.int64)
    tcol = table[time_column].combine_chunks()
    if not pa.types.is_timestamp(tcol.type):
        raise TypeError(f"column {time_column!r} must be a timestamp column, got {tcol.type}")
    unit, tz = tcol.type.unit, tcol.type.tz
    ticks = tcol.cast(pa.int64()).fill_null(F.NAT).to_numpy(zero_copy_only=False).astype(np.int64, copy=False)
    key = np.where(ticks == F.NAT, I64_MAX, ticks)
    order = None
    n = len(codes)
    grouped = n < 2 or bool(((codes[1:] > codes[:-1]) | ((codes[1:] == codes[:-1]) & (key[1:] >= key[:-1]))).all())
    if not grouped or (n and codes.min() < 0):
        order = np.lexsort((key, codes))
        if n and codes.min() < 0:
            order = order[int((codes < 0).sum()):]      # null symbols sort first: drop them
        codes, ticks = codes[order], ticks[order]
    n_sym = len(enc.dictionary)
    counts = np.bincount(codes, minlength=n_sym) if len(codes) else np.zeros(n_sym, np.int64)
    off = np.zeros(n_sym + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    return SymbolGroups(enc.dictionary, off, order, np.ascontiguousarray(ticks), unit, tz)


def _ordered(a: np.ndarray, g: SymbolGroups) -> np.ndarray:
    return a if g.order is None else np.ascontiguousarray(a[g.order])


def _symbol_column(g: SymbolGroups, out_off: np.ndarray) -> pa.Array:
    idx = np.repeat(np.arange(len(g.symbols), dtype=np.int32), np.diff(out_off))
    return pa.DictionaryArray.from_arrays(pa.array(idx), g.symbols)


def _stamps(ticks: np.ndarray, g: SymbolGroups) -> pa.Array:
    return pa.array(ticks).cast(pa.timestamp(g.unit, g.tz))


def _dev():
    import torch
    if not torch.cuda.is_available():
        raise IvxError("no CUDA device: iv_interpolation_b200 has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _t(a: np.ndarray, dev):
    return _to_device(np.ascontiguousarray(a), dev)


# ---------------------------------------------------------------------------------------------------------------
# Task 1
# ---------------------------------------------------------------------------------------------------------------
def _task1_device(table: pa.Table, min_points: int, launch=None, dev=None):
    """Group, upload and run kernel A; returns everything the table builders need (device results not yet downloaded)."""
    from . import device as D
    for c in ("date",) + F.DROPNA_COLS:
        if c not in table.column_names:
            raise KeyError(c)
    g = group_by_symbol(table, "date")
    dev = dev if dev is not None else _dev()
    num_names = [c for c in F.FILL_COLS if c in table.column_names and c != "symbol" and
                 (pa.types.is_floating(table[c].type) or pa.types.is_integer(table[c].type) or pa.types.is_decimal(table[c].type))]
    str_names = [c for c in F.FILL_COLS if c in table.column_names and c != "symbol" and c not in num_names]
    if len(num_names) > 8 or len(str_names) > 3:
        raise IvxError("at most 8 numeric and 3 text forward-fill columns besides symbol")
    icols = [_t(_ordered(_f64(table[c]), g), dev) for c in F.INTERP_COLS]
    fcols = [_t(_ordered(_f64(table[c]), g), dev) for c in num_names]
    masks = [None]                                      # symbol: never null inside a group
    for c in str_names:
        col = table[c]
        masks.append(None if col.null_count == 0 else _t(_ordered(pc.is_valid(col).combine_chunks().to_numpy(zero_copy_only=False), g).astype(np.uint8), dev))
    step = 60 * F.ticks_per_second(g.unit)
    res = (launch or D.interp_linear_ffill)(_t(g.off, dev), _t(g.ticks, dev), icols, fcols, masks, step=step,
                                            max_span=30 * 86400 * F.ticks_per_second(g.unit), min_points=min_points,
                                            want_keep=True, want_is_interpolated=False)
    return g, num_names, str_names, res


def interpolate_table(table: pa.Table, method: str = "linear", min_points: int = 10, launch=None, dev=None) -> Tuple[pa.Table, Dict[str, dict]]:
    """Every symbol of ``table`` through Task 1.  Returns the interpolated rows (reference column order + ``is_interpolated``)
    and a per-symbol status dictionary in the reference's words (batch_processor.py:99-126)."""
    if method not in ("linear", "index", "values"):
        raise IvxError(f"method {method!r} is not available (the reference returns None for it too: core.py:83-85)")
    g, num_names, str_names, res = _task1_device(table, min_points, launch, dev)
    host = d2h_many([res.sym_status, res.out_off, res.ts, res.keep] + list(res.interp) + list(res.fill) + list(res.idx))
    status, out_off, o_ts, o_keep = host[:4]
    o_interp = host[4:7]
    o_fill = host[7:7 + len(num_names)]
    o_idx = host[7 + len(num_names):]
    cols = {"date": _stamps(o_ts, g), "symbol": _symbol_column(g, out_off)}
    for c, a in zip(F.INTERP_COLS, o_interp):
        cols[c] = pa.array(a)
    for c, a in zip(num_names, o_fill):
        cols[c] = pa.array(a)
    for k, c in enumerate(str_names):
        ix = o_idx[k + 1].astype(np.int64)
        if g.order is not None:
            ix = np.where(ix >= 0, g.order[np.maximum(ix, 0)], -1)
        cols[c] = pc.take(table[c], pa.array(ix, mask=ix < 0))
    names = [c for c in TASK1_ORDER if c in cols]
    out = pa.table([cols[c] for c in names], names=names)
    out = out.append_column("is_interpolated", pa.array(np.zeros(len(o_ts), dtype=bool)))     # core.py:71: always False
    keep = o_keep.view(bool)
    if not keep.all():                                   # +-inf among the interpolated columns: dropna cuts interior rows
        out = out.filter(pa.array(keep))
        out_off = np.concatenate([[0], np.cumsum(keep)])[out_off]
    statuses = {}
    in_rows, n_out = np.diff(g.off), np.diff(out_off)
    for i, s in enumerate(g.symbols.to_pylist()):
        if status[i] == SYM_OK and n_out[i] > 0:
            statuses[s] = {"symbol": s, "status": "success", "input_rows": int(in_rows[i]), "output_rows": int(n_out[i])}
        else:
            statuses[s] = {"symbol": s, "status": "skipped", "reason": "Interpolation failed or insufficient data",
                           "detail": _STATUS_REASON.get(int(status[i]), "No valid data after interpolation")}
    return out, statuses


# ---------------------------------------------------------------------------------------------------------------
# Task 2
# ---------------------------------------------------------------------------------------------------------------
def parse_frequency(freq: str) -> int:
    """candle core.py:27-34."""
    if freq.endswith("min"):
        return int(freq[:-3])
    if freq.endswith("m"):
        return int(freq[:-1])
    raise ValueError(f"Unsupported frequency: {freq}")


def reconstruct_table(table: pa.Table, target_frequency: str = "5min", launch=None, dev=None) -> Tuple[pa.Table, Dict[str, dict]]:
    """Every symbol of a 1-minute OHLCV table through Task 2; columns of candle core.py:97-99."""
    from . import device as D
    n = parse_frequency(target_frequency)
    if n < 1:
        raise ValueError(f"Unsupported frequency: {target_frequency}")
    missing = [c for c in F.CANDLE_REQUIRED + ("symbol",) if c not in table.column_names]
    if missing:
        raise KeyError(f"Missing required columns: {missing}")
    g = group_by_symbol(table, "timestamp")
    dev = dev if dev is not None else _dev()
    cols = [_t(_ordered(_f64(table[c]), g), dev) for c in ("open", "high", "low", "close", "volume")]
    res = (launch or D.ohlcv_reduce)(_t(g.off, dev), _t(g.ticks, dev), *cols, bucket_ticks=n * 60 * F.ticks_per_second(g.unit),
                                     min_count=n, check_len=True, want_index=False, trim=True)
    status, flags, out_off, o_ts, oo, oh, ol, oc, ov = d2h_many([res.sym_status, res.sym_flags, res.out_off, res.ts, res.open, res.high,
                                                                 res.low, res.close, res.volume])
    m = len(o_ts)
    out = pa.table({"symbol": _symbol_column(g, out_off), "timestamp": _stamps(o_ts, g), "open": pa.array(oo), "high": pa.array(oh),
                    "low": pa.array(ol), "close": pa.array(oc), "volume": pa.array(ov),
                    "frequency": pa.DictionaryArray.from_arrays(pa.array(np.zeros(m, np.int32)), pa.array([target_frequency])),
                    "source_candles": pa.array(np.full(m, n, np.int64)),
                    "created_at": pa.array(np.full(m, np.datetime64(datetime.now(), "us")))})
    statuses = {}
    in_rows, n_out = np.diff(g.off), np.diff(out_off)
    for i, s in enumerate(g.symbols.to_pylist()):
        fl = int(flags[i])
        if fl & 7:                                       # candle core.py:229-230
            statuses[s] = {"symbol": s, "status": "error", "error": "Invalid input candle data"}
        elif status[i] != SYM_OK:                        # :235-236
            statuses[s] = {"symbol": s, "status": "skipped", "reason": "Reconstruction failed"}
        elif n_out[i] == 0 or fl & 56:                   # :239-240 (validate(empty) is False)
            statuses[s] = {"symbol": s, "status": "error", "error": "Invalid reconstructed candle data"}
        else:
            statuses[s] = {"symbol": s, "status": "success", "input_candles": int(in_rows[i]), "output_candles": int(n_out[i])}
    return out, statuses


# ---------------------------------------------------------------------------------------------------------------
# pipeline: Task 1 -> bridge -> Task 2 on the device
# ---------------------------------------------------------------------------------------------------------------
def pipeline_table(table: pa.Table, *, min_points: int = 10, strategy: str = "spread_simulation", target_frequency: str = "5min",
                   seed: int = 0, want: Sequence[str] = ("candles",)) -> Dict[str, object]:
    """Task 1 -> data bridge -> Task 2 for every symbol of a Task-1 table, chained on the device (the reference writes each
    stage to PostgreSQL and reads it back: complete_pipeline.py:740-831).  ``want`` names the stages whose tables are brought
    back: "interpolated", "minute", "candles".  The bridge's random numbers come from the device's counter generator keyed
    by ``seed`` (the reference draws from an unseeded generator; seeded parity runs go through the DataFrame classes)."""
    from . import device as D
    n = parse_frequency(target_frequency)
    g, num_names, str_names, a = _task1_device(table, min_points)
    if "volume" not in num_names:
        raise KeyError("volume")
    S, vol = a.interp[1], a.fill[num_names.index("volume")]
    bplan = D.bridge_plan(a.out_off, S, vol)
    # the draws are made inside the bridge kernel (the elements random_tape(seed, stream 0 / 1) would hold) and the stamps copied there
    m = D.bridge_ohlcv_philox(a.out_off, S, vol, bplan, strategy, seed=seed, stream_u=0, stream_e=1, ts=a.ts, want_row=False)
    mts = m.ts
    step = 60 * F.ticks_per_second(g.unit)
    c = D.ohlcv_reduce(m.out_off, mts, m.open, m.high, m.low, m.close, m.volume, bucket_ticks=n * step, min_count=n, check_len=True)
    out: Dict[str, object] = {"symbols": g.symbols}
    status, a_off, m_off, c_off = d2h_many([a.sym_status, a.out_off, m.out_off, c.out_off])
    out["counts"] = {"source_rows": np.diff(g.off), "interpolated": np.diff(a_off), "minute": np.diff(m_off), "candles": np.diff(c_off),
                     "status": status}
    if "interpolated" in want:
        h = d2h_many([a.ts] + list(a.interp) + list(a.fill))
        cols = {"date": _stamps(h[0], g), "symbol": _symbol_column(g, a_off)}
        for name, arr in zip(F.INTERP_COLS + tuple(num_names), h[1:]):
            cols[name] = pa.array(arr)
        names = [x for x in TASK1_ORDER if x in cols]
        out["interpolated"] = pa.table([cols[x] for x in names], names=names)
    if "minute" in want:
        h = d2h_many([mts, m.open, m.high, m.low, m.close, m.volume, m.source_price])
        out["minute"] = pa.table({"symbol": _symbol_column(g, m_off), "timestamp": _stamps(h[0], g), "open": pa.array(h[1]), "high": pa.array(h[2]),
                                  "low": pa.array(h[3]), "close": pa.array(h[4]), "volume": pa.array(h[5]), "source_price": pa.array(h[6])})
    if "candles" in want:
        h = d2h_many([c.ts, c.open, c.high, c.low, c.close, c.volume])
        mm = len(h[0])
        out["candles"] = pa.table({"symbol": _symbol_column(g, c_off), "timestamp": _stamps(h[0], g), "open": pa.array(h[1]), "high": pa.array(h[2]),
                                   "low": pa.array(h[3]), "close": pa.array(h[4]), "volume": pa.array(h[5]),
                                   "frequency": pa.DictionaryArray.from_arrays(pa.array(np.zeros(mm, np.int32)), pa.array([target_frequency])),
                                   "source_candles": pa.array(np.full(mm, n, np.int64))})
    return out


def bridge_table(table: pa.Table, strategy: str = "spread_simulation", seed: int = 0) -> pa.Table:
    """Interpolated rows (``symbol, date|timestamp, underlying_price, volume``) -> synthetic 1-minute candles, device tapes."""
    from . import device as D
    tcol = "timestamp" if "timestamp" in table.column_names else "date"
    g = group_by_symbol(table, tcol)
    dev = _dev()
    price = next((c for c in ("underlying_price", "mark_price", "index_price") if c in table.column_names), None)
    if price is None:
        raise ValueError("No suitable price column found")
    S = _t(_ordered(_f64(table[price]), g), dev)
    vol = _t(_ordered(_f64(table["volume"]), g), dev) if "volume" in table.column_names else S.new_zeros(S.numel())
    off, ts = _t(g.off, dev), _t(g.ticks, dev)
    bplan = D.bridge_plan(off, S, vol)
    m = D.bridge_ohlcv_philox(off, S, vol, bplan, strategy, seed=seed, stream_u=0, stream_e=1, ts=ts, want_row=False)
    h = d2h_many([m.out_off, m.ts, m.open, m.high, m.low, m.close, m.volume, m.source_price])
    return pa.table({"symbol": _symbol_column(g, h[0]), "timestamp": _stamps(h[1], g), "open": pa.array(h[2]), "high": pa.array(h[3]),
                     "low": pa.array(h[4]), "close": pa.array(h[5]), "volume": pa.array(h[6]), "source_price": pa.array(h[7])})
