"""Columnar in/out at the reference's persisted precision.

The reference keeps its data in PostgreSQL and spends most of its wall time turning frames into row tuples
(``iterrows`` + ``execute_values``, reference src/interpolation/batch_processor.py:167-187,
src/candle_reconstruction/core.py:284-320).  This module is the DB-free counterpart of that boundary: Arrow
tables / Parquet files with the reference's column names and the DECIMAL precisions of its DDL
(reference src/database/schema.py:20-52, src/database/candle_schema.py:22-49), so that results can be compared
at the precision the reference would have stored.  Host-side plumbing only; nothing here is on the hot path.
"""
from __future__ import annotations

from typing import Dict, Iterable, Mapping

import numpy as np
import pandas as pd
import pyarrow as pa
import pyarrow.parquet as pq

# column -> (precision, scale): schema.py:28-46
INTERPOLATED_DECIMALS = {
    "iv": (12, 8), "underlying_price": (15, 4), "time_to_maturity": (12, 8), "strike": (15, 4), "interest_rate": (8, 6),
    "delta": (12, 8), "gamma": (12, 8), "theta": (12, 8), "vega": (12, 8), "rho": (12, 8),
    "mark_price": (15, 4), "index_price": (15, 4), "volume": (20, 8), "quote_volume": (20, 8),
}
INTERPOLATED_ORDER = ("symbol", "date", "record_time", "iv", "underlying_price", "time_to_maturity", "strike", "callput",
                      "interest_rate", "delta", "gamma", "theta", "vega", "rho", "mark_price", "index_price", "volume",
                      "quote_volume", "is_interpolated", "batch_id")
# candle_schema.py:30-39
CANDLE_DECIMALS = {"open": (15, 4), "high": (15, 4), "low": (15, 4), "close": (15, 4), "volume": (20, 8)}
CANDLE_ORDER = ("symbol", "timestamp", "open", "high", "low", "close", "volume", "frequency", "source_candles", "batch_id",
                "created_at")


def quantize(values, scale: int) -> np.ndarray:
    """Round half to even at ``scale`` decimals — what ``DECIMAL(p, s)`` keeps of a float8 (PostgreSQL numeric
    rounding of a double's shortest decimal representation differs from this only on exact binary ties)."""
    a = np.asarray(values, dtype=np.float64)
    return np.round(a, scale)


def _decimal_array(col: pd.Series, precision: int, scale: int) -> pa.Array:
    a = pa.array(quantize(col.to_numpy(dtype=np.float64, na_value=np.nan), scale), from_pandas=True)
    return a.cast(pa.decimal128(precision, scale), safe=False)


def frame_to_table(df: pd.DataFrame, decimals: Mapping[str, tuple], order: Iterable[str]) -> pa.Table:
    cols, names = [], []
    for c in order:
        if c not in df.columns:
            continue
        if c in decimals:
            cols.append(_decimal_array(df[c], *decimals[c]))
        else:
            cols.append(pa.array(df[c], from_pandas=True))
        names.append(c)
    return pa.table(cols, names=names)


def interpolated_to_table(frames: Mapping[str, pd.DataFrame]) -> pa.Table:
    """All contracts of a batch as one table with the columns of ``interpolated_trading_tickers`` (schema.py:20-52)."""
    parts = [frame_to_table(f, INTERPOLATED_DECIMALS, INTERPOLATED_ORDER) for f in frames.values() if f is not None and len(f)]
    return pa.concat_tables(parts, promote_options="default") if parts else pa.table({})


def candles_to_table(frames: Mapping[str, pd.DataFrame]) -> pa.Table:
    """``reconstructed_candles`` (candle_schema.py:22-49)."""
    parts = [frame_to_table(f, CANDLE_DECIMALS, CANDLE_ORDER) for f in frames.values() if f is not None and len(f)]
    return pa.concat_tables(parts, promote_options="default") if parts else pa.table({})


def table_to_db_precision(table: pa.Table, decimals: Mapping[str, tuple]) -> pa.Table:
    """An Arrow table with every column named in ``decimals`` rounded (half to even) and stored as ``DECIMAL(p, s)`` —
    NaN becomes NULL, as the reference's INSERT does with ``None`` — and everything else untouched: what the reference's
    tables would hold (schema.py:20-52, candle_schema.py:22-49), without going through pandas."""
    import pyarrow.compute as pc
    for i, name in enumerate(table.column_names):
        if name not in decimals:
            continue
        p, s = decimals[name]
        col = table.column(i)
        if not pa.types.is_floating(col.type):
            continue
        col = pc.if_else(pc.is_nan(col), pa.scalar(None, pa.float64()), col)
        col = pc.round(col, ndigits=s, round_mode="half_to_even").cast(pa.decimal128(p, s), safe=False)
        table = table.set_column(i, name, col)
    return table


def write_parquet(table: pa.Table, path: str) -> None:
    pq.write_table(table, path, compression="zstd")


def read_symbol_frames(path: str, time_column: str) -> Dict[str, pd.DataFrame]:
    """A Parquet file with a ``symbol`` column -> {symbol: frame ordered by time}: the DB-free stand-in for the
    reference's per-symbol ``SELECT ... WHERE symbol = %s ORDER BY date`` (batch_processor.py:76-87,
    candle_reconstruction/core.py:215-223).  DECIMAL columns come back as float64, as ``pd.read_sql`` coerces them."""
    t = pq.read_table(path)
    cols = {}
    for name in t.column_names:
        c = t.column(name)
        cols[name] = c.cast(pa.float64()) if pa.types.is_decimal(c.type) else c
    df = pa.table(cols).to_pandas()
    return {str(sym): g.sort_values(time_column, kind="stable").reset_index(drop=True)
            for sym, g in df.groupby("symbol", sort=False)}
