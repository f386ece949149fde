"""InterpolatedToOHLCVConverter — the reference's data bridge (src/data_bridge/ohlcv_converter.py:16-418) on the B200.

Same class name, constructor and method names as the reference; the candle arithmetic runs in libivx.so
(``ivx_bridge_plan`` + ``ivx_bridge_ohlcv``, include/ivx.h "Data bridge").  No CPU fallback.

Randomness.  The reference calls ``np.random.uniform / exponential / normal`` row by row on numpy's GLOBAL generator.
This class takes its draws from the same place in the same order (one ``random_sample`` per uniform or exponential,
``standard_normal`` for the trend-following noise), so after ``np.random.seed(s)`` it returns the frames the
reference returns after ``np.random.seed(s)`` — bit for bit — and leaves the generator in the same state.  Pass
``rng=np.random.RandomState(...)`` to use a private stream instead.

Database I/O (``get_interpolated_symbols``, ``convert_symbol_to_ohlcv``'s SELECT, ``_save_ohlcv_candles``,
``convert_batch``: reference :31-136, :420-520) is out of scope; ``convert_frames`` is the DB-free
batch entry with the reference's per-symbol status dictionaries (:113-135).
"""
from __future__ import annotations

import logging
import math
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import pandas as pd

logger = logging.getLogger(__name__)

PRICE_PRIORITY = ["underlying_price", "mark_price", "index_price"]            # reference :192
_METHOD_NAME = {"spread_simulation": "spread_simulation", "price_as_midpoint": "price_midpoint",
                "trend_following": "trend_following", "simple_spread": "simple_spread"}
_Q_REASONS = ((1, "High < Low found"), (2, "High < Open/Close found"), (4, "Low > Open/Close found"),
              (8, "Unrealistic spreads detected"), (16, "Negative or zero prices found"))
_U_PER_ROW = {"spread_simulation": 5, "price_as_midpoint": 2, "trend_following": 0, "simple_spread": 1, "pipeline": 4}


def canonical_strategy(name: str) -> str:
    """The reference's dispatch (:158-166): unknown strings — including the 'price_midpoint' its own config comment
    suggests (config.py:60) — take the simple-spread fallback."""
    return name if name in ("spread_simulation", "price_as_midpoint", "trend_following", "pipeline") else "simple_spread"


def draw_tapes(rs, strategy: str, ok: np.ndarray, need_vol: np.ndarray):
    """Draw (u, e, z) for the emitted rows of one or more symbols from ``rs`` (``np.random`` or a ``RandomState``) in the
    reference's call order.  ``ok``: rows that emit a candle; ``need_vol``: rows among them whose volume is drawn."""
    nv = need_vol[ok]
    n_rows, n_e = int(ok.sum()), int(nv.sum())
    if strategy == "trend_following":
        if n_e == 0:
            return np.zeros(0), np.zeros(0), np.asarray(rs.standard_normal(n_rows), dtype=np.float64).reshape(-1)
        z, e = np.empty(n_rows), []
        for i, need in enumerate(nv):              # normals and exponentials interleave in the generator: call by call
            z[i] = rs.standard_normal()
            if need:
                e.append(rs.standard_exponential())
        return np.zeros(0), np.asarray(e, dtype=np.float64), z
    k = _U_PER_ROW[strategy]
    per_row = k + nv.astype(np.int64)
    stream = np.asarray(rs.random_sample(int(per_row.sum())), dtype=np.float64).reshape(-1)
    start = np.cumsum(per_row) - per_row
    u = stream[(start[:, None] + np.arange(k)[None, :]).reshape(-1)] if n_rows else np.zeros(0)
    # legacy standard_exponential is -log(1 - u) through the C library's log; math.log is that function
    e = np.asarray([-math.log(1.0 - float(stream[s + k])) for s in start[nv]], dtype=np.float64)
    return u, e, np.zeros(0)


class InterpolatedToOHLCVConverter:
    """Convert Task 1 interpolated rows to Task 2 OHLCV candles (B200 implementation)."""

    def __init__(self, db_manager=None, config=None, rng=None, device=None):
        self.db_manager = db_manager
        self.config = config
        bridge = getattr(config, "data_bridge", None)
        self.conversion_strategy = getattr(bridge, "conversion_strategy", "spread_simulation")      # :26
        self.spread_method = getattr(bridge, "spread_method", "adaptive")
        self.quality_checks = getattr(bridge, "enable_quality_checks", True)
        # without a config: the reference's defaults (config.py:69-75); with one: whatever it holds — a None there makes the
        # reference's spread strategy fail on ``None.get`` (:210) and return None, and so does this class
        self._spread_parameters = {} if bridge is None else getattr(bridge, "spread_parameters", None)
        self._rng = rng
        self._device = device

    # -- the reference's per-frame entry (:138-188) ------------------------------------------------------
    def _generate_ohlcv_from_interpolated(self, df: pd.DataFrame) -> Optional[pd.DataFrame]:
        if df.empty:
            return None
        try:
            out = self._generate_many([df], self.conversion_strategy)[0]
            return out["frame"]
        except Exception as e:                                                    # :186-188
            logger.error(f"OHLCV generation failed: {e}")
            return None

    # complete_pipeline.py:473-510 — same bridge, fixed 0.1 % spread, no clean-up; returns an EMPTY frame when nothing is emitted
    def _generate_ohlcv_candles(self, data: pd.DataFrame) -> pd.DataFrame:
        res = self._generate_many([data], "pipeline")[0]["frame"]
        return res if res is not None else pd.DataFrame()

    def _select_price_column(self, df: pd.DataFrame) -> str:
        """Reference :189-207."""
        for col in PRICE_PRIORITY:
            if col in df.columns:
                non_null_count = df[col].notna().sum()
                if non_null_count > len(df) * 0.8:
                    logger.debug(f"Using {col} as primary price source ({non_null_count}/{len(df)} valid)")
                    return col
        for col in PRICE_PRIORITY:
            if col in df.columns:
                logger.warning(f"Using {col} as fallback price source (limited data)")
                return col
        raise ValueError("No suitable price column found")

    def _validate_ohlcv_quality(self, ohlcv_df: pd.DataFrame) -> Dict:
        """Reference :371-402 on a frame: the five checks run on the device (``ivx_bridge_quality``), the verdict is the
        first failing one in the reference's order; any error inside is reported the way the reference reports it."""
        try:
            import torch

            from .. import device as D
            dev = torch.device(self._device) if self._device is not None else torch.device("cuda")
            cols = [torch.from_numpy(np.array(ohlcv_df[k].to_numpy(dtype=np.float64), copy=True)).to(dev)
                    for k in ("open", "high", "low", "close", "source_price")]
            return self._quality_verdict(D.bridge_quality(*cols))
        except Exception as e:
            return {"valid": False, "reason": f"Validation error: {str(e)}"}

    @staticmethod
    def _quality_verdict(flags: int) -> Dict:
        for bit, reason in _Q_REASONS:
            if flags & bit:
                return {"valid": False, "reason": reason}
        return {"valid": True, "reason": "All quality checks passed"}

    # -- DB-free batch entry: frames in, frames + the reference's status dictionaries out (:113-135) --------------------
    def convert_frames(self, frames: Sequence[pd.DataFrame], symbols: Optional[Sequence[str]] = None) -> List[Dict]:
        t0 = time.time()
        symbols = list(symbols) if symbols is not None else [
            (str(f["symbol"].iloc[0]) if len(f) and "symbol" in f.columns else f"#{i}") for i, f in enumerate(frames)]
        results: List[Dict] = []
        try:
            gen = self._generate_many(frames, self.conversion_strategy)
        except Exception as e:                                                    # :127-136
            return [{"symbol": s, "status": "error", "error": str(e), "processing_time": time.time() - t0} for s in symbols]
        dt = time.time() - t0
        for sym, f, g in zip(symbols, frames, gen):
            if f.empty:
                results.append({"symbol": sym, "status": "skipped", "reason": "No interpolated data found"})       # :92
            elif g["frame"] is None or g["frame"].empty:
                results.append({"symbol": sym, "status": "error", "error": "OHLCV conversion failed"})            # :98
            elif self.quality_checks and not g["quality"]["valid"]:
                results.append({"symbol": sym, "status": "error", "error": f"Quality check failed: {g['quality']['reason']}"})
            else:
                results.append({"symbol": sym, "status": "success", "input_points": len(f), "output_candles": len(g["frame"]),
                                "processing_time": dt / max(1, len(frames)), "conversion_method": self.conversion_strategy,
                                "ohlcv": g["frame"]})
        return results

    # -- the reference's database entries (:31-136, :420-520): out of scope, and saying so beats an AttributeError ------------
    def _no_database(self, name: str):
        raise NotImplementedError(f"InterpolatedToOHLCVConverter.{name} is PostgreSQL I/O in the reference (DESIGN.md section 7); "
                                  "read the frames yourself (iv_interpolation_b200.columnar) and call convert_frames(frames)")

    def get_interpolated_symbols(self, batch_id=None):
        self._no_database("get_interpolated_symbols")

    def convert_symbol_to_ohlcv(self, symbol: str, batch_id=None):
        self._no_database("convert_symbol_to_ohlcv")

    def convert_batch(self, symbols=None, batch_id=None):
        self._no_database("convert_batch")

    def _spread_kwargs(self, strategy: str) -> Dict:
        if strategy != "spread_simulation":
            return {}
        cfg = self._spread_parameters                       # reference :207-211; raises on None like the reference does
        return {"base_spread_percent": cfg.get("base_spread_percent", 0.002), "volatility_factor": cfg.get("volatility_factor", 1.5)}

    # -- one launch for any number of frames -------------------------------------------------------------------------
    def _uniform(self, frames: Sequence[pd.DataFrame], ts_col: str) -> bool:
        """Frames sharing one column layout with a tz-naive datetime64 stamp column, float64 prices/volume and a symbol column:
        they are concatenated once and the result is handed out as slices of one frame (see hostpath.py)."""
        from ..hostpath import FrameSlices
        if isinstance(frames, FrameSlices):                # one frame already (its slices share the layout by construction)
            live = [frames.big] if len(frames.big) else []
        else:
            live = [f for f in frames if not f.empty]
            if len(live) < 2:
                return False
        if not live:
            return False
        first = live[0]
        cols, dtypes = list(first.columns), list(first.dtypes)
        if "symbol" not in cols or ts_col not in cols or "volume" not in cols:
            return False
        dt = first[ts_col].dtype
        if not (isinstance(dt, np.dtype) and dt.kind == "M"):
            return False
        if any(c in cols and first[c].dtype != np.float64 for c in PRICE_PRIORITY + ["volume"]) or not any(c in cols for c in PRICE_PRIORITY):
            return False
        return all(list(f.columns) == cols and list(f.dtypes) == dtypes for f in live[1:])

    def _generate_many(self, frames: Sequence[pd.DataFrame], strategy_name: str, whole_out: Optional[Dict] = None) -> List[Dict]:
        """``frames``: a list of frames or a ``hostpath.FrameSlices`` (then nothing is concatenated); ``whole_out`` (a dict)
        receives the one frame the results are slices of, when there is one: ``frame``, ``off``."""
        import torch

        from .. import device as D
        from ..hostpath import FrameSlices, d2h_many

        strategy = canonical_strategy(strategy_name)
        rs = self._rng if self._rng is not None else np.random
        dev = torch.device(self._device) if self._device is not None else torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        pipeline = strategy == "pipeline"
        ts_col = "date" if pipeline else "timestamp"
        if isinstance(frames, FrameSlices):
            lens, off = np.diff(frames.off), frames.off
        else:
            lens = np.fromiter((len(f) for f in frames), dtype=np.int64, count=len(frames))
            off = np.zeros(len(frames) + 1, np.int64)
            np.cumsum(lens, out=off[1:])
        uniform = self._uniform(frames, ts_col)
        big = None
        if uniform:
            big = frames.big if isinstance(frames, FrameSlices) else pd.concat([f for f in frames if not f.empty], ignore_index=True)
            nz = np.flatnonzero(lens)
            starts = off[:-1][nz]
            if pipeline:                                   # complete_pipeline.py:478: a or b or c on floats (NaN truthy, 0.0 not)
                u_, m_, i_ = (big[c].to_numpy(dtype=np.float64) for c in PRICE_PRIORITY)
                base = np.where(u_ != 0, u_, np.where(m_ != 0, m_, i_))
            else:                                          # _select_price_column per frame (:189-207), all frames at once
                have = [c for c in PRICE_PRIORITY if c in big.columns]
                counts = np.stack([np.add.reduceat(big[c].notna().to_numpy(dtype=np.int64), starts) for c in have])
                good = counts > lens[nz][None, :] * 0.8
                choice = np.where(good.any(axis=0), good.argmax(axis=0), 0)
                if not good.any(axis=0).all():
                    logger.warning(f"Using {have[0]} as fallback price source (limited data)")
                per_row = np.repeat(choice, lens[nz])
                base = big[have[0]].to_numpy(dtype=np.float64).copy()
                for k in range(1, len(have)):
                    sel = per_row == k
                    if sel.any():
                        base[sel] = big[have[k]].to_numpy(dtype=np.float64)[sel]
            vol = big["volume"].to_numpy(dtype=np.float64)
        else:
            bases, vols = [], []
            for f in frames:
                if f.empty:
                    bases.append(np.zeros(0)); vols.append(np.zeros(0))
                    continue
                if pipeline:
                    u_, m_, i_ = (f[c].to_numpy(dtype=np.float64) for c in PRICE_PRIORITY)
                    b = np.where(u_ != 0, u_, np.where(m_ != 0, m_, i_))
                else:
                    col = self._select_price_column(f)         # raises ValueError -> caller's except, like :144-148
                    b = f[col].to_numpy(dtype=np.float64)
                v = f["volume"].to_numpy(dtype=np.float64) if "volume" in f.columns else np.zeros(len(f))   # row.get('volume', 0)
                bases.append(b); vols.append(v)
            base = np.concatenate(bases) if bases else np.zeros(0)
            vol = np.concatenate(vols) if vols else np.zeros(0)
        with np.errstate(invalid="ignore"):
            ok = ~(np.isnan(base) | (base <= 0))
            need = ok & (np.isnan(vol) | (vol <= 0))
        u, e, z = draw_tapes(rs, strategy, ok, need)       # symbol after symbol, row after row: the reference's order
        from ..hostpath import _to_device
        t = lambda a: _to_device(np.ascontiguousarray(a), dev)
        d_off, d_base, d_vol = t(off), t(base), t(vol)
        plan = D.bridge_plan(d_off, d_base, d_vol)
        if plan.total != int(ok.sum()) or plan.total_e != int(need.sum()):
            raise RuntimeError("bridge plan disagrees with the host's row classification")
        res = D.bridge_ohlcv(d_off, d_base, d_vol, plan, strategy, u=t(u) if len(u) else None, e=t(e), z=t(z) if len(z) else None,
                             **self._spread_kwargs(strategy))
        names = ("open", "high", "low", "close", "volume", "source_price")
        host = d2h_many([res.out_off, res.row, res.sym_quality] + [getattr(res, k) for k in names])
        out_off, row, quality = host[:3]
        cols = dict(zip(names, host[3:]))
        if uniform:
            got = self._assemble_uniform(frames, big, ts_col, out_off, row, cols, quality, pipeline, strategy, whole_out)
            if got is not None:
                return got
        out: List[Dict] = []
        for i, f in enumerate(frames):
            a, b_ = int(out_off[i]), int(out_off[i + 1])
            if b_ == a:
                out.append({"frame": None, "quality": self._quality_verdict(0)})          # :170-171 no candles -> None
                continue
            src = row[a:b_] - off[i]
            data = {"symbol": f["symbol"].to_numpy()[src] if "symbol" in f.columns else None}
            if data["symbol"] is None:
                raise KeyError("symbol")
            data["timestamp"] = pd.to_datetime(f[ts_col]).to_numpy()[src]
            for k in names:
                data[k] = cols[k][a:b_]
            if not pipeline:
                data["conversion_method"] = _METHOD_NAME[strategy]
            data["is_synthetic"] = True
            frame = pd.DataFrame(data)
            if not pipeline:                               # _cleanup_ohlcv_data :404-418
                frame = frame.dropna(subset=["open", "high", "low", "close"])
                frame = frame.sort_values("timestamp")
            out.append({"frame": frame, "quality": self._quality_verdict(int(quality[i]))})
        return out

    def _assemble_uniform(self, frames, big, ts_col, out_off, row, cols, quality, pipeline, strategy, whole_out=None):
        """All candles as ONE frame, handed out as per-symbol slices.  Returns None (-> per-frame assembly) when the clean-up
        of :404-418 would actually change something: a NaN among the OHLC columns or stamps out of order inside a symbol."""
        nf = len(frames)
        n_out = np.diff(out_off)
        ts_all = big[ts_col].to_numpy()
        every = len(row) == len(ts_all)                    # every row emitted a candle: `row` is the identity, nothing to gather
        ts = ts_all if every else ts_all[row]
        if not pipeline:
            if any(np.isnan(cols[k]).any() for k in ("open", "high", "low", "close")):
                return None
            if len(ts) > 1:
                fid = np.repeat(np.arange(nf), n_out)
                if ((ts[1:] < ts[:-1]) & (fid[1:] == fid[:-1])).any():
                    return None
        data = {"symbol": big["symbol"].array if every else big["symbol"].array.take(row), "timestamp": ts}
        for k in ("open", "high", "low", "close", "volume", "source_price"):
            data[k] = cols[k]
        if not pipeline:
            data["conversion_method"] = _METHOD_NAME[strategy]
        data["is_synthetic"] = True
        local = np.arange(len(row), dtype=np.int64) - np.repeat(out_off[:-1], n_out)
        frame = pd.DataFrame(data, index=pd.Index(local, copy=False))
        if whole_out is not None:
            whole_out.update(frame=frame, off=np.asarray(out_off, dtype=np.int64))
        out: List[Dict] = []
        for i in range(nf):
            a, b_ = int(out_off[i]), int(out_off[i + 1])
            if b_ == a:
                out.append({"frame": None, "quality": self._quality_verdict(0)})
            else:
                out.append({"frame": frame.iloc[a:b_], "quality": self._quality_verdict(int(quality[i]))})
        return out
