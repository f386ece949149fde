"""CandleReconstructor — same class and methods as reference src/candle_reconstruction/core.py:11-171,
computed by the sm_100a kernels of libivx.so (kernel B).  No CPU fallback.

Error contract (reference core.py:47-49,55-57,64-66,104-106): ``reconstruct_symbol_candles`` never
raises; it returns ``None`` for empty/insufficient/invalid input and an EMPTY frame when no bucket
is complete.  The constructor raises ``ValueError`` for an unsupported frequency (core.py:34).
"""
from __future__ import annotations

import logging
from datetime import datetime
from typing import List, Optional, Sequence

import numpy as np
import pandas as pd

from .. import frames as F

logger = logging.getLogger(__name__)


class CandleReconstructor:
    """Reconstruct N-minute candles from 1-minute OHLCV data (B200 implementation).

    open = first non-null open, high = max, low = min, close = last non-null close,
    volume = (Kahan) sum; only buckets holding at least N source rows are kept.
    """

    def __init__(self, target_frequency: str = "5min"):
        self.target_frequency = target_frequency
        self.frequency_minutes = self._parse_frequency(target_frequency)

    def _parse_frequency(self, freq: str) -> int:
        """Parse frequency string to minutes (reference core.py:27-34)."""
        if freq.endswith("min"):
            return int(freq[:-3])
        elif freq.endswith("m"):
            return int(freq[:-1])
        else:
            raise ValueError(f"Unsupported frequency: {freq}")

    # -- single symbol: the reference's entry point --------------------------------------------------
    def reconstruct_symbol_candles(self, minute_data: pd.DataFrame) -> Optional[pd.DataFrame]:
        if minute_data.empty:                                                    # core.py:47-49
            logger.warning("No minute data provided")
            return None
        try:
            return self.reconstruct_batch([minute_data])[0]
        except Exception as e:                                                   # core.py:104-106
            logger.error(f"Candle reconstruction failed: {e}")
            return None

    # -- many symbols in one launch ---------------------------------------------------------------------
    def reconstruct_batch(self, frames: Sequence[pd.DataFrame], presorted: bool = False,
                          check_len: bool = True, flags_out: Optional[list] = None, whole_out: Optional[dict] = None,
                          slices: bool = True) -> List[Optional[pd.DataFrame]]:
        """``[reconstruct_symbol_candles(f) for f in frames]`` with a single kernel launch.

        ``presorted=True, check_len=False`` is the contract of the reference's second entry,
        CompleteOptimizedPipeline._aggregate_to_5min (complete_pipeline.py:685-710): rows already
        ordered by the SQL query, no length early-out.  ``frames`` may be a ``hostpath.FrameSlices`` (no concat);
        ``whole_out`` (a dict) receives the one frame the results are slices of, when there is one (``frame``, ``off``, ``ok``,
        ``slots``), and ``slices=False`` then leaves the per-symbol results to the caller (all ``None``).
        """
        import torch
        from .. import device as D
        from .._lib import SYM_OK, IvxError, lib

        lib()
        if not torch.cuda.is_available():
            raise IvxError("no CUDA device: iv_interpolation_b200 has no CPU path")
        n = self.frequency_minutes
        results: List[Optional[pd.DataFrame]] = [None] * len(frames)
        if flags_out is not None:           # per frame: IVX_FLAG_* bits of the launch (validate_candle_data of input and output), None = not run
            flags_out[:] = [None] * len(frames)
        from .. import hostpath as H
        whole = isinstance(frames, H.FrameSlices) and len(frames) > 1 and bool((np.diff(frames.off) > 0).all())
        live = list(range(len(frames))) if whole else [i for i, df in enumerate(frames) if not df.empty]
        batch = frames if whole else [frames[i] for i in live]
        if n >= 1 and len(live) > 1 and H.task2_uniform(batch):
            wo = {} if whole_out is not None else None
            got = H.reconstruct_uniform(batch, n, self.target_frequency, datetime.now(), presorted=presorted, check_len=check_len,
                                        whole_out=wo, slices=slices or whole_out is None)
            from .._lib import FLAG_UNSORTED
            if got is not None and not (presorted and any(f & FLAG_UNSORTED for f in got[1])):
                if whole_out is not None:
                    whole_out.update(wo, slots=live)
                outs, fl = got
                for k, i in enumerate(live):
                    results[i] = outs[k]
                    if flags_out is not None:
                        flags_out[i] = fl[k]
                return results
        inputs, srcs, slots = [], [], []
        for i, df in enumerate(frames):
            if df.empty:
                continue
            missing = [c for c in F.CANDLE_REQUIRED if c not in df.columns]
            if missing:                                                          # core.py:53-57
                logger.error(f"Missing required columns: {missing}")
                continue
            try:
                src = df.reset_index(drop=True)
                inp = F.candle_frame_to_arrays(src)
            except Exception as e:
                logger.error(f"Candle reconstruction failed: {e}")
                continue
            if check_len and inp.n < n:                                          # core.py:64-66
                logger.warning(f"Insufficient data: {inp.n} rows, need at least {n}")
                continue
            if n < 1:                     # dt.floor('0min') raises in the reference -> None
                logger.error("Candle reconstruction failed: non-positive frequency")
                continue
            order = None
            key = np.where(inp.ts == F.NAT, np.iinfo(np.int64).max, inp.ts)
            if len(key) > 1 and np.any(key[1:] < key[:-1]):
                if not presorted:                                                # core.py:60
                    order = np.argsort(key, kind="stable")
                else:
                    # complete_pipeline.py:685-710 does not sort, but its groupby still MERGES the rows of a bucket wherever they
                    # stand, first/last in input order: rows are brought together bucket by bucket, input order kept inside one
                    width = n * 60 * F.ticks_per_second(inp.unit)
                    bucket = np.where(inp.ts == F.NAT, np.iinfo(np.int64).max, (inp.ts + inp.wall_shift) // width)
                    order = np.argsort(bucket, kind="stable")
            inputs.append((inp, order)); srcs.append(src); slots.append(i)
        if not inputs:
            return results
        groups = {}
        for j, (inp, _) in enumerate(inputs):
            groups.setdefault(inp.unit, []).append(j)
        dev = torch.device("cuda")
        created_at = datetime.now()                                              # core.py:94
        for unit, members in groups.items():
            lens = [inputs[j][0].n for j in members]
            off = np.zeros(len(members) + 1, np.int64)
            np.cumsum(lens, out=off[1:])

            def cat(get):
                parts = []
                for j in members:
                    inp, order = inputs[j]
                    a = get(inp)
                    parts.append(a if order is None else a[order])
                return np.ascontiguousarray(np.concatenate(parts))

            ts = torch.from_numpy(cat(lambda p: np.where(p.ts == F.NAT, F.NAT, p.ts + p.wall_shift) if p.wall_shift else p.ts)).to(dev)
            cols = [torch.from_numpy(cat(lambda p, k=k: p.ohlcv[k])).to(dev) for k in range(5)]
            any_null = any(inputs[j][0].sym_valid is not None for j in members)
            sv = None
            if any_null:
                sv = torch.from_numpy(cat(lambda p: (p.sym_valid if p.sym_valid is not None
                                                     else np.ones(p.n, bool))).astype(np.uint8)).to(dev)
            res = D.ohlcv_reduce(torch.from_numpy(off).to(dev), ts, *cols,
                                 bucket_ticks=n * 60 * F.ticks_per_second(unit), min_count=n,
                                 check_len=check_len, sym_valid=sv, want_index=True)
            status = res.sym_status.cpu().numpy()
            out_off = res.out_off.cpu().numpy()
            goff = res.sym_group_off.cpu().numpy()
            if flags_out is not None:
                fl = res.sym_flags.cpu().numpy()
                for g, j in enumerate(members):
                    flags_out[slots[j]] = int(fl[g])
            o_ts = res.ts.cpu().numpy()
            o_cols = [t.cpu().numpy() for t in (res.open, res.high, res.low, res.close, res.volume)]
            o_sym = res.sym_row.cpu().numpy()
            o_grp = res.group.cpu().numpy()
            for g, j in enumerate(members):
                if status[g] != SYM_OK:
                    continue
                inp, order = inputs[j]
                a, b = int(out_off[g]), int(out_off[g + 1])
                base = int(off[g])
                sym_row = o_sym[a:b]
                loc = np.where(sym_row >= 0, sym_row - base, 0)
                if order is not None:
                    loc = order[loc]
                sym_row = np.where(sym_row >= 0, loc, -1)
                results[slots[j]] = F.candle_arrays_to_frame(
                    srcs[j], inp, o_ts[a:b], [c[a:b] for c in o_cols], sym_row, o_grp[a:b] - goff[g],
                    self.target_frequency, n, created_at)
                logger.debug(f"Reconstructed {inp.n} 1-min candles → {b - a} {self.target_frequency} candles")
        return results

    def aggregate_to_5min(self, data: pd.DataFrame) -> pd.DataFrame:
        """The reference's second entry, CompleteOptimizedPipeline._aggregate_to_5min
        (complete_pipeline.py:685-710): N=5, rows taken in the given order, seven columns back."""
        five = CandleReconstructor("5min")
        out = five.reconstruct_batch([data], presorted=True, check_len=False)[0]
        if out is None:
            return pd.DataFrame(columns=["timestamp", "open", "high", "low", "close", "volume", "symbol"])
        return out[["timestamp", "open", "high", "low", "close", "volume", "symbol"]]

    # -- validation / statistics (reference core.py:108-171) ---------------------------------------------------
    def validate_candle_data(self, df: pd.DataFrame) -> bool:
        """Validate OHLCV candle data integrity on the device (flag reductions of kernel B)."""
        if df.empty:
            return False
        try:
            import torch
            from .. import device as D
            from .._lib import FLAG_IN_BAD_OHLC, FLAG_IN_NAN, FLAG_IN_NEG_VOLUME

            dev = torch.device("cuda")
            cols = [torch.from_numpy(np.array(df[c].to_numpy(dtype=np.float64), dtype=np.float64, copy=True)).to(dev)
                    for c in ("open", "high", "low", "close", "volume")]
            m = len(df)
            off = torch.tensor([0, m], dtype=torch.int64, device=dev)
            ts = torch.arange(m, dtype=torch.int64, device=dev)        # one row per bucket: flags only
            res = D.ohlcv_reduce(off, ts, *cols, bucket_ticks=1, min_count=1, check_len=False, trim=False)
            flags = int(res.sym_flags.cpu().numpy()[0])
            if flags & FLAG_IN_BAD_OHLC:
                logger.warning("Found candles with invalid OHLC relationships")
                return False
            if flags & FLAG_IN_NEG_VOLUME:
                logger.warning("Found candles with negative volume")
                return False
            if flags & FLAG_IN_NAN:
                logger.warning("Found null values in OHLC columns")
                return False
            logger.debug("Candle data validation passed")
            return True
        except Exception as e:                                                   # core.py:142-144
            logger.error(f"Validation failed: {e}")
            return False

    def get_reconstruction_stats(self, original_data: pd.DataFrame, reconstructed_data: pd.DataFrame) -> dict:
        """Statistics about the reconstruction (reference core.py:146-171); host-side, as the survey
        allows: a handful of scalar reductions over frames already on the host."""
        if original_data.empty or reconstructed_data.empty:
            return {}
        try:
            original_timespan = original_data["timestamp"].max() - original_data["timestamp"].min()
            reconstructed_timespan = reconstructed_data["timestamp"].max() - reconstructed_data["timestamp"].min()
            vol_in = original_data["volume"].sum()
            vol_out = reconstructed_data["volume"].sum()
            return {
                "original_candles": len(original_data),
                "reconstructed_candles": len(reconstructed_data),
                "compression_ratio": len(original_data) / len(reconstructed_data) if len(reconstructed_data) > 0 else 0,
                "original_timespan": original_timespan,
                "reconstructed_timespan": reconstructed_timespan,
                "coverage_ratio": reconstructed_timespan / original_timespan if original_timespan.total_seconds() > 0 else 0,
                "total_volume_original": vol_in,
                "total_volume_reconstructed": vol_out,
                "volume_preservation": abs(1 - vol_out / vol_in) if vol_in > 0 else 1,
            }
        except Exception as e:
            logger.error(f"Failed to calculate stats: {e}")
            return {}
