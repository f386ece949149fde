"""Host-side DataFrame <-> structure-of-arrays conversion for the drop-in classes.

Pure pandas/numpy; no device code and no arithmetic of the hot path happens here.  The
column contracts follow the reference:

  * Task 1 input  (reference src/interpolation/batch_processor.py:76-83): ``symbol, date,
    iv, underlying_price, time_to_maturity, strike, callput, interest_rate, mark_price,
    index_price, volume, quote_volume, record_time``;
  * Task 1 output (reference src/interpolation/core.py:54-74): ``date`` first, the input's
    other columns in input order, then ``is_interpolated``;
  * Task 2 input/output (reference src/candle_reconstruction/core.py:53,97-99).

Timestamps travel as int64 ticks in the column's own datetime64 unit (pandas 3 infers
``us``, pandas 2 ``ns``); NaT is INT64_MIN, which the kernels treat as "no stamp".
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd

NAT = np.iinfo(np.int64).min

INTERP_COLS = ("iv", "underlying_price", "time_to_maturity")           # core.py:58
FILL_COLS = ("symbol", "strike", "callput", "interest_rate", "mark_price",
             "index_price", "volume", "quote_volume", "record_time")    # core.py:64-65
DROPNA_COLS = ("symbol", "iv", "underlying_price", "time_to_maturity")  # core.py:74
CANDLE_REQUIRED = ("timestamp", "open", "high", "low", "close", "volume")  # candle core.py:53
CANDLE_OUT_ORDER = ("symbol", "timestamp", "open", "high", "low", "close", "volume",
                    "frequency", "source_candles", "created_at")        # candle core.py:97-98

_TICKS_PER_SECOND = {"s": 1, "ms": 10 ** 3, "us": 10 ** 6, "ns": 10 ** 9}


def ticks_per_second(unit: str) -> int:
    return _TICKS_PER_SECOND[unit]


def datetime_to_ticks(col: pd.Series) -> Tuple[np.ndarray, str, object]:
    """``pd.to_datetime(col)`` (core.py:33, candle core.py:61) -> (int64 ticks, unit, tz).

    Raises whatever ``pd.to_datetime`` raises; the callers turn that into ``None``.
    """
    dt = pd.to_datetime(col)
    dtype = dt.dtype
    tz = getattr(dtype, "tz", None)
    unit = getattr(dtype, "unit", None) or np.datetime_data(dtype)[0]
    if unit not in _TICKS_PER_SECOND:
        raise TypeError(f"unsupported datetime unit {unit!r}")
    if tz is not None:
        vals = dt.dt.tz_convert("UTC").dt.tz_localize(None).to_numpy(dtype=f"datetime64[{unit}]")
    else:
        vals = dt.to_numpy(dtype=f"datetime64[{unit}]")
    return vals.view(np.int64).copy(), unit, tz


def ticks_to_datetime(ticks: np.ndarray, unit: str, tz=None) -> pd.Series:
    ser = pd.Series(np.asarray(ticks, dtype=np.int64).view(f"datetime64[{unit}]"))
    if tz is not None:
        ser = ser.dt.tz_localize("UTC").dt.tz_convert(tz)
    return ser


def _is_plain_numeric(dtype) -> bool:
    return isinstance(dtype, np.dtype) and dtype.kind in "fiub"


@dataclass
class InterpInput:
    """One contract's Task-1 input as arrays (input row order, unsorted allowed)."""
    n: int
    unit: str
    tz: object
    date: np.ndarray                                    # int64[n]
    interp: List[np.ndarray]                            # 3 x float64[n]
    fill_num_names: List[str] = field(default_factory=list)
    fill_num: List[np.ndarray] = field(default_factory=list)       # float64[n] each
    fill_idx_names: List[str] = field(default_factory=list)        # string / non-numeric fill cols
    fill_idx_valid: List[np.ndarray] = field(default_factory=list)  # bool[n] each (non-null)
    columns: List[str] = field(default_factory=list)    # input column order (incl. date)
    fill_int_dtype: Dict[str, object] = field(default_factory=dict)  # integer forward-fill columns travelling as float64

    def grid_is_dense(self) -> bool:
        """Every grid point between min(date) and max(date) carries a source row: the left merge (core.py:54-55) then
        introduces no NaN and pandas keeps an integer column's dtype; otherwise the column comes back float64."""
        d = self.date[self.date != NAT]
        if len(d) == 0:
            return False
        t0, step = int(d.min()), self.step
        on = (d - t0) % step == 0
        return int(np.unique((d[on] - t0) // step).size) == (int(d.max()) - t0) // step + 1

    @property
    def step(self) -> int:                               # core.py:45 freq='1min'
        return 60 * _TICKS_PER_SECOND[self.unit]

    @property
    def max_span(self) -> int:                           # core.py:37 timedelta(days=30)
        return 30 * 86400 * _TICKS_PER_SECOND[self.unit]


def interp_frame_to_arrays(df: pd.DataFrame) -> InterpInput:
    """Split a Task-1 frame into arrays.  Raises KeyError/TypeError/ValueError on frames the
    reference would also fail on (missing ``date``/``symbol``/interp columns, unparsable dates,
    non-numeric interp columns); the caller maps that to ``None`` (core.py:83-85)."""
    for c in ("date",) + DROPNA_COLS:                    # core.py:32 and :74 raise KeyError
        if c not in df.columns:
            raise KeyError(c)
    date, unit, tz = datetime_to_ticks(df["date"])
    interp = []
    for c in INTERP_COLS:
        if not _is_plain_numeric(df[c].dtype):
            raise TypeError(f"column {c!r} has non-numeric dtype {df[c].dtype}")
        interp.append(np.ascontiguousarray(df[c].to_numpy(dtype=np.float64)))
    out = InterpInput(n=len(df), unit=unit, tz=tz, date=date, interp=interp, columns=list(df.columns))
    for c in FILL_COLS:
        if c not in df.columns:
            continue
        col = df[c]
        if _is_plain_numeric(col.dtype) and col.dtype.kind == "f":
            out.fill_num_names.append(c)
            out.fill_num.append(np.ascontiguousarray(col.to_numpy(dtype=np.float64)))
        elif _is_plain_numeric(col.dtype) and col.dtype.kind in "iu" and (len(col) == 0 or int(np.abs(col.to_numpy()).max()) < 2 ** 53):
            # integer strike / volume ...: exact in float64, and float64 is what the reference returns once the left merge has
            # introduced a NaN (interp_arrays_to_frame restores the dtype on a dense grid, where it does not)
            out.fill_num_names.append(c)
            out.fill_num.append(np.ascontiguousarray(col.to_numpy(dtype=np.float64)))
            out.fill_int_dtype[c] = col.dtype
        else:
            out.fill_idx_names.append(c)
            out.fill_idx_valid.append(np.ascontiguousarray(col.notna().to_numpy(dtype=bool)))
    return out


def interp_arrays_to_frame(
    src: pd.DataFrame,
    inp: InterpInput,
    out_date: np.ndarray,
    out_interp: List[np.ndarray],
    out_fill_num: List[np.ndarray],
    out_fill_idx: List[np.ndarray],
    out_src_row: np.ndarray,
    index: np.ndarray,
) -> pd.DataFrame:
    """Assemble the Task-1 output frame (column order of core.py:54-55 merge + :71)."""
    m = len(out_date)
    cols: Dict[str, object] = {}
    cols["date"] = ticks_to_datetime(out_date, inp.unit, inp.tz).to_numpy() if inp.tz is None \
        else ticks_to_datetime(out_date, inp.unit, inp.tz).array        # .array: no index alignment against the position labels
    interp_map = dict(zip(INTERP_COLS, out_interp))
    num_map = dict(zip(inp.fill_num_names, out_fill_num))
    idx_map = dict(zip(inp.fill_idx_names, out_fill_idx))
    for c in inp.columns:
        if c == "date":
            continue
        if c in interp_map:
            cols[c] = interp_map[c]
        elif c in num_map:
            cols[c] = num_map[c]
            if c in inp.fill_int_dtype and inp.grid_is_dense():
                cols[c] = np.asarray(num_map[c]).astype(inp.fill_int_dtype[c])
        elif c in idx_map:
            cols[c] = _take_with_null(src[c], idx_map[c])
        else:                                           # extra column: merged but never filled
            cols[c] = _take_with_null(src[c], out_src_row)
    out = pd.DataFrame(cols, index=pd.Index(index) if m else None)
    out["is_interpolated"] = np.zeros(m, dtype=bool)   # core.py:71 — always False after ffill
    return out


def _take_with_null(col: pd.Series, idx: np.ndarray):
    """col.iloc[idx] with idx == -1 -> missing (what the left merge leaves there)."""
    idx = np.asarray(idx, dtype=np.int64)
    if len(idx) and idx.min() >= 0:
        return col.take(idx).reset_index(drop=True).to_numpy() if _is_plain_numeric(col.dtype) \
            else col.take(idx).reset_index(drop=True).array
    taken = col.take(np.maximum(idx, 0)).reset_index(drop=True)
    if _is_plain_numeric(taken.dtype) and taken.dtype.kind != "f":
        taken = taken.astype(np.float64)
    taken = taken.mask(pd.Series(idx < 0))
    return taken.array


# ---------------------------------------------------------------------------------------
# Task 2
# ---------------------------------------------------------------------------------------
@dataclass
class CandleInput:
    n: int
    unit: str
    tz: object
    ts: np.ndarray                   # int64[n]
    ohlcv: List[np.ndarray]          # 5 x float64[n]
    sym_valid: Optional[np.ndarray]  # bool[n] or None when every symbol is non-null
    wall_shift: int = 0              # tz-aware stamps: ticks to add to ``ts`` (UTC) to get the local wall clock the reference floors


def candle_frame_to_arrays(df: pd.DataFrame) -> CandleInput:
    """Raises KeyError when ``symbol`` is missing (candle core.py:79 -> :104-106 None)."""
    if "symbol" not in df.columns:
        raise KeyError("symbol")
    ts, unit, tz = datetime_to_ticks(df["timestamp"])
    shift = 0
    if tz is not None and len(ts):
        # ``dt.floor`` on a tz-aware column (candle core.py:70) floors the LOCAL wall clock, not UTC: the two differ when the
        # bucket width does not divide the UTC offset (60-minute candles in +05:30).  With one offset for the whole frame the
        # wall clock is UTC + offset and the buckets are formed there; a frame that straddles an offset change (DST) is refused.
        wall = pd.to_datetime(df["timestamp"]).dt.tz_localize(None).to_numpy(dtype=f"datetime64[{unit}]").view(np.int64)
        ok = ts != NAT
        off = np.unique(wall[ok] - ts[ok])
        if len(off) > 1:
            raise ValueError("the UTC offset changes inside the frame (DST): bucket boundaries are ambiguous")
        shift = int(off[0]) if len(off) else 0
    ohlcv = []
    for c in ("open", "high", "low", "close", "volume"):
        if not _is_plain_numeric(df[c].dtype):
            raise TypeError(f"column {c!r} has non-numeric dtype {df[c].dtype}")
        ohlcv.append(np.ascontiguousarray(df[c].to_numpy(dtype=np.float64)))
    valid = df["symbol"].notna().to_numpy(dtype=bool)
    return CandleInput(n=len(df), unit=unit, tz=tz, ts=ts, ohlcv=ohlcv,
                       sym_valid=None if valid.all() else np.ascontiguousarray(valid), wall_shift=shift)


def candle_arrays_to_frame(
    src: pd.DataFrame,
    inp: CandleInput,
    out_ts: np.ndarray,
    out_ohlcv: List[np.ndarray],
    out_sym_row: np.ndarray,
    index: np.ndarray,
    target_frequency: str,
    source_candles: int,
    created_at,
) -> pd.DataFrame:
    """candle core.py:82-99: rename, metadata, column order; index is NOT reset."""
    m = len(out_ts)
    ts = ticks_to_datetime(np.asarray(out_ts, dtype=np.int64) - inp.wall_shift, inp.unit, inp.tz)   # bucket starts back from wall clock to UTC
    data = {
        "symbol": _take_with_null(src["symbol"], out_sym_row),
        "timestamp": ts.to_numpy() if inp.tz is None else ts.array,      # .array: no index alignment against the group labels
        "open": out_ohlcv[0], "high": out_ohlcv[1], "low": out_ohlcv[2],
        "close": out_ohlcv[3], "volume": out_ohlcv[4],
    }
    out = pd.DataFrame(data, index=pd.Index(np.asarray(index, dtype=np.int64)) if m else None)
    out["frequency"] = target_frequency                  # core.py:92
    out["source_candles"] = source_candles               # core.py:93
    out["created_at"] = created_at                       # core.py:94
    return out[list(CANDLE_OUT_ORDER)]
