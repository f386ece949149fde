"""DataFrame-level faces of the numpy oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Same constructor/method names as the reference classes so the parity tests read like the
reference's call sites (src/interpolation/batch_processor.py:21-24,97;
src/candle_reconstruction/core.py:180,229-248).  The frame <-> array conversion is the
product's host code (iv_interpolation_b200/frames.py); the arithmetic is oracle/np_oracle.py.
"""
from __future__ import annotations

from datetime import datetime
from typing import Optional

import numpy as np
import pandas as pd

from iv_interpolation_b200 import frames as F
from oracle import np_oracle as O

LINEAR_METHODS = ("linear", "index", "values")


class OracleIVInterpolator:
    """Restates IVInterpolator (reference src/interpolation/core.py:9-85)."""

    def __init__(self, method: str = "linear", min_points: int = 10):
        self.method = method
        self.min_points = min_points

    def interpolate_symbol(self, symbol_data: pd.DataFrame) -> Optional[pd.DataFrame]:
        if len(symbol_data) < self.min_points:
            return None
        if self.method not in LINEAR_METHODS:
            return None
        try:
            inp = F.interp_frame_to_arrays(symbol_data)
        except Exception:
            return None
        res = self._run(inp)
        if res["status"] != O.SYM_OK:
            return None
        sidx = dict(zip(res["str_names"], res["str_src"]))
        return F.interp_arrays_to_frame(
            symbol_data.reset_index(drop=True), inp, res["date"], list(res["interp"]), list(res["fill"]),
            [sidx[c] for c in inp.fill_idx_names], res["src_row"], res["index"],
        )

    def _run(self, inp: F.InterpInput):
        # symbol's validity must be first: it drives dropna (core.py:74)
        names = list(inp.fill_idx_names)
        valid = list(inp.fill_idx_valid)
        extra_symbol = False
        if "symbol" in names:
            i = names.index("symbol")
            names.insert(0, names.pop(i))
            valid.insert(0, valid.pop(i))
        else:
            # symbol is a float column (all-NaN / numeric ids): validity from the numeric copy
            j = inp.fill_num_names.index("symbol")
            names.insert(0, "symbol")
            valid.insert(0, ~np.isnan(inp.fill_num[j]))
            extra_symbol = True
        res = O.interp_linear_ffill(
            inp.date, inp.interp, inp.fill_num, valid,
            step=inp.step, max_span=inp.max_span, min_points=self.min_points,
        )
        if res["status"] != O.SYM_OK:
            return res
        res["str_names"] = names
        if extra_symbol:
            res["str_names"] = names[1:]
            res["str_src"] = res["str_src"][1:]
        return res


class OracleCandleReconstructor:
    """Restates CandleReconstructor (reference src/candle_reconstruction/core.py:11-171)."""

    def __init__(self, target_frequency: str = "5min"):
        self.target_frequency = target_frequency
        self.frequency_minutes = self._parse_frequency(target_frequency)

    def _parse_frequency(self, freq: str) -> int:            # candle core.py:27-34
        if freq.endswith("min"):
            return int(freq[:-3])
        elif freq.endswith("m"):
            return int(freq[:-1])
        raise ValueError(f"Unsupported frequency: {freq}")

    def reconstruct_symbol_candles(self, minute_data: pd.DataFrame) -> Optional[pd.DataFrame]:
        if minute_data.empty:                                  # :47-49
            return None
        if any(c not in minute_data.columns for c in F.CANDLE_REQUIRED):   # :53-57
            return None
        try:
            inp = F.candle_frame_to_arrays(minute_data)
            n = self.frequency_minutes
            res = O.ohlcv_reduce(
                np.where(inp.ts == F.NAT, F.NAT, inp.ts + inp.wall_shift) if inp.wall_shift else inp.ts, *inp.ohlcv, bucket_ticks=n * 60 * F.ticks_per_second(inp.unit),
                min_count=n, sym_valid=inp.sym_valid,
            )
        except Exception:
            return None
        if res["status"] != O.SYM_OK:
            return None
        return F.candle_arrays_to_frame(
            minute_data.reset_index(drop=True), inp, res["ts"],
            [res[k] for k in ("open", "high", "low", "close", "volume")],
            res["sym_row"], res["group_index"], self.target_frequency, n, datetime.now(),
        )

    def validate_candle_data(self, df: pd.DataFrame) -> bool:  # :108-144
        if df.empty:
            return False
        try:
            return O.validate_candles(*(df[c].to_numpy(dtype=np.float64)
                                        for c in ("open", "high", "low", "close", "volume")))
        except Exception:
            return False
