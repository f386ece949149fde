"""Batched host path of the drop-in classes: many frames -> ONE set of columns -> one launch -> ONE result frame, sliced.

The per-frame path of ``interpolation/core.py`` / ``candle_reconstruction/core.py`` converts every frame on its own and
builds every output frame column by column in Python.  With thousands of contracts per call (BASELINE.json configs[0]:
2 400 contracts x 25 hourly rows) that host work, not the kernel, is the whole cost.  Frames that share the reference's
plain schema (reference src/interpolation/batch_processor.py:76-83: three string columns, ``date``, nine float64 columns;
reference src/candle_reconstruction/core.py:215-223 for Task 2) take this path instead:

  * ``pd.concat`` once, every column extracted once, sortedness checked for all frames in one vectorised pass;
  * one plan + exec launch (kernel A) or one reduce launch (kernel B) for the whole batch;
  * every result column leaves the device through pinned host memory with ONE synchronisation (``d2h_many``);
  * one big result frame, handed out as per-symbol slices (``iloc``), the reference's index labels kept.

Anything else (extra columns, integer or object-typed numeric columns, tz-aware stamps, mixed schemas) goes through the
general per-frame path, which defines the semantics; the two are held equal by the golden tests.  No arithmetic of the hot
path happens here and nothing falls back to the CPU: the launches go through ``device.py`` -> libivx.so.
"""
from __future__ import annotations

import time
import warnings
from typing import Callable, List, Optional, Sequence

import numpy as np
import pandas as pd

from . import frames as F

I64_MAX = np.iinfo(np.int64).max
TASK1_STR = ("symbol", "callput", "record_time")
TASK1_NUM = F.INTERP_COLS + ("strike", "interest_rate", "mark_price", "index_price", "volume", "quote_volume")
TASK1_COLS = frozenset(("date",) + TASK1_STR + TASK1_NUM)
TASK2_NUM = ("open", "high", "low", "close", "volume")


class FrameSlices:
    """A batch of per-symbol frames that ARE consecutive row ranges of one frame: ``self[i] == big.iloc[off[i]:off[i+1]]``.

    A stage of ``batch.PipelineBatch`` hands its whole result frame to the next stage in this form, so the next stage skips
    the per-symbol column selection and the ``pd.concat`` that would rebuild the very frame the slices were cut from
    (0.5-2 ms of pandas per symbol and stage: at 2 400 contracts that, not the launches, was the cost of the pipeline).
    Built only by this package from frames it has just produced itself."""

    def __init__(self, big: pd.DataFrame, off: np.ndarray):
        off = np.asarray(off, dtype=np.int64)
        if len(off) < 1 or off[0] != 0 or off[-1] != len(big) or (len(off) > 1 and (np.diff(off) < 0).any()):
            raise ValueError("FrameSlices: offsets do not tile the frame")
        self.big, self.off = big, off

    def __len__(self) -> int:
        return len(self.off) - 1

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return self.big.iloc[int(self.off[i]):int(self.off[i + 1])]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _batch_of(frames):
    """(nf, lens, off, big) of a batch: the frame itself for ``FrameSlices``, one ``pd.concat`` for a list of frames."""
    if isinstance(frames, FrameSlices):
        return len(frames), np.diff(frames.off), frames.off, frames.big
    nf = len(frames)
    lens = np.fromiter((len(f) for f in frames), dtype=np.int64, count=nf)
    off = np.zeros(nf + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    return nf, lens, off, (frames[0] if nf == 1 else pd.concat(frames, ignore_index=True))


def d2h_many(tensors: Sequence[object]) -> List[Optional[np.ndarray]]:
    """Device tensors -> numpy arrays through pinned host memory (torch's caching host allocator: the blocks are reused
    across calls), all copies enqueued before the single synchronisation.  ``None`` entries pass through."""
    import torch
    hosts = []
    any_cuda = False
    for t in tensors:
        if t is None:
            hosts.append(None)
        elif not t.is_cuda:
            hosts.append(t)
        else:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            hosts.append(h)
            any_cuda = True
    if any_cuda:
        torch.cuda.current_stream().synchronize()
    return [None if h is None else h.numpy() for h in hosts]


def _to_device(a: np.ndarray, dev):
    """numpy -> device tensor (pandas hands out read-only views under copy-on-write; the copy to the device never writes)."""
    import torch
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return torch.from_numpy(a).to(dev)


def _naive_datetime(dtype) -> Optional[str]:
    if isinstance(dtype, np.dtype) and dtype.kind == "M":
        unit = np.datetime_data(dtype)[0]
        return unit if unit in ("s", "ms", "us", "ns") else None
    return None


def task1_uniform(frames: Sequence[pd.DataFrame]) -> bool:
    """Every frame has exactly the reference's 13 Task-1 columns in one common order, float64 numerics, one tz-naive
    datetime64 unit for ``date`` and non-numeric string columns."""
    if not len(frames):
        return False
    if isinstance(frames, FrameSlices):
        frames = [frames.big]
    first = frames[0]
    cols = list(first.columns)
    if len(cols) != len(TASK1_COLS) or set(cols) != TASK1_COLS:
        return False
    dtypes = list(first.dtypes)
    if _naive_datetime(first["date"].dtype) is None:
        return False
    for c in TASK1_NUM:
        if first[c].dtype != np.float64:
            return False
    for c in TASK1_STR:
        dt = first[c].dtype
        if isinstance(dt, np.dtype) and dt.kind in "fiubcmM":
            return False
    for f in frames[1:]:
        if list(f.columns) != cols or list(f.dtypes) != dtypes:
            return False
    return True


def _take_fill(values, idx: np.ndarray):
    """values[idx] with idx == -1 -> missing, dtype kept (what the left merge + ffill leave in a string column)."""
    return values.take(idx, allow_fill=bool(len(idx)) and bool(idx.min() < 0))


def _tick(timing: Optional[dict], key: str, t0: float) -> float:
    t1 = time.perf_counter()
    if timing is not None:
        timing[key] = timing.get(key, 0.0) + (t1 - t0)
    return t1


def interpolate_uniform(frames: Sequence[pd.DataFrame], min_points: int, launch: Optional[Callable] = None,
                        device=None, timing: Optional[dict] = None, extra: Optional[dict] = None,
                        whole_out: Optional[dict] = None) -> List[Optional[pd.DataFrame]]:
    """``[IVInterpolator.interpolate_symbol(f) for f in frames]`` for frames that satisfy ``task1_uniform`` and hold at
    least ``min_points`` rows each (reference core.py:26-28 is the caller's business)."""
    import torch
    from . import device as D
    from ._lib import SYM_OK

    launch = launch or D.interp_linear_ffill
    dev = device if device is not None else torch.device("cuda")
    tk = time.perf_counter()
    nf, lens, off, big = _batch_of(frames)
    unit = _naive_datetime(big["date"].dtype)
    date = big["date"].to_numpy().view(np.int64)
    order = None
    if len(date) > 1:                                                   # rows travel in time order (core.py:32), NaT last
        key = np.where(date == F.NAT, I64_MAX, date)
        fid = np.repeat(np.arange(nf, dtype=np.int64), lens)
        bad = (key[1:] < key[:-1]) & (fid[1:] == fid[:-1])
        if bad.any():
            order = np.lexsort((key, fid))                               # stable inside a frame
    pick = (lambda a: np.ascontiguousarray(a)) if order is None else (lambda a: np.ascontiguousarray(a[order]))
    t = lambda a: _to_device(a, dev)
    step = 60 * F.ticks_per_second(unit)
    masks = []
    for c in TASK1_STR:                                                 # symbol first: it takes part in dropna (core.py:74)
        m = big[c].notna().to_numpy(dtype=bool)
        masks.append(None if m.all() else t(pick(m).astype(np.uint8)))
    tk = _tick(timing, "host_concat_and_columns_s", tk)
    res = launch(t(off), t(pick(date)), [t(pick(big[c].to_numpy())) for c in F.INTERP_COLS],
                 [t(pick(big[c].to_numpy())) for c in TASK1_NUM[3:]], masks,
                 step=step, max_span=30 * 86400 * F.ticks_per_second(unit), min_points=min_points,
                 want_src_row=False, want_pos=True, want_keep=True, want_is_interpolated=False)
    host = d2h_many([res.sym_status, res.sym_flags, res.out_off, res.ts, res.pos, res.keep] + list(res.interp) + list(res.fill) + list(res.idx))
    tk = _tick(timing, "h2d_launch_d2h_s", tk)
    status, flags, out_off, o_ts, o_pos, o_keep = host[:6]
    o_interp, o_fill, o_idx = host[6:9], host[9:15], host[15:18]
    keep = o_keep.view(bool)
    all_kept = bool(keep.all())
    if not all_kept:                                                    # only with +-inf among the interpolated columns
        sel = np.flatnonzero(keep)
        cut = lambda a: a[sel]
        kept_before = np.concatenate([[0], np.cumsum(keep)])
        out_off = kept_before[out_off]
    else:
        cut = lambda a: a
    data = {"date": cut(o_ts).view(f"datetime64[{unit}]")}
    interp_map = dict(zip(F.INTERP_COLS, o_interp))
    fill_map = dict(zip(TASK1_NUM[3:], o_fill))
    idx_map = dict(zip(TASK1_STR, o_idx))
    for c in big.columns:
        if c == "date":
            continue
        if c in interp_map:
            data[c] = cut(interp_map[c])
        elif c in fill_map:
            data[c] = cut(fill_map[c])
        else:
            ix = cut(idx_map[c]).astype(np.int64)
            if order is not None:
                ix = np.where(ix >= 0, order[np.maximum(ix, 0)], -1)
            data[c] = _take_fill(big[c].array, ix)
    m = len(data["date"])
    out = pd.DataFrame(data, index=pd.Index(cut(o_pos).astype(np.int64), copy=False), copy=False)
    out["is_interpolated"] = np.zeros(m, dtype=bool)                    # core.py:71 — always False after the forward fill
    for name, value in (extra or {}).items():                           # e.g. batch_id (batch_processor.py:105): once, not per symbol
        out[name] = value
    results: List[Optional[pd.DataFrame]] = [None] * nf
    for g in range(nf):
        if status[g] != SYM_OK:
            continue
        a, b = int(out_off[g]), int(out_off[g + 1])
        if b > a:                                                       # core.py:76-78: empty after dropna -> None
            results[g] = out.iloc[a:b]
    if whole_out is not None:                                           # the frame the results are slices of (see FrameSlices)
        whole_out.update(frame=out, off=np.asarray(out_off, dtype=np.int64), ok=np.asarray(status) == SYM_OK)
    _tick(timing, "host_result_frames_s", tk)
    return results


def task2_uniform(frames: Sequence[pd.DataFrame]) -> bool:
    """Task-2 frames with one common column order, float64 OHLCV, one tz-naive datetime64 unit and no null symbol."""
    if not len(frames):
        return False
    if isinstance(frames, FrameSlices):
        frames = [frames.big]
    first = frames[0]
    cols = list(first.columns)
    if "symbol" not in cols or any(c not in cols for c in F.CANDLE_REQUIRED):
        return False
    if _naive_datetime(first["timestamp"].dtype) is None or any(first[c].dtype != np.float64 for c in TASK2_NUM):
        return False
    dt = first["symbol"].dtype
    if isinstance(dt, np.dtype) and dt.kind in "fiubcmM":
        return False
    dtypes = list(first.dtypes)
    for f in frames[1:]:
        if list(f.columns) != cols or list(f.dtypes) != dtypes:
            return False
    return True


def reconstruct_uniform(frames: Sequence[pd.DataFrame], n: int, target_frequency: str, created_at, *, presorted: bool,
                        check_len: bool, launch: Optional[Callable] = None, device=None, whole_out: Optional[dict] = None,
                        slices: bool = True):
    """``[reconstruct_symbol_candles(f) for f in frames]`` for non-empty ``task2_uniform`` frames; returns (results, flags)
    or ``None`` when a symbol column holds nulls (the per-frame path handles first-non-null semantics)."""
    import torch
    from . import device as D
    from ._lib import SYM_OK

    launch = launch or D.ohlcv_reduce
    dev = device if device is not None else torch.device("cuda")
    nf, lens, off, big = _batch_of(frames)
    if big["symbol"].isna().any():
        return None
    unit = _naive_datetime(big["timestamp"].dtype)
    ts = big["timestamp"].to_numpy().view(np.int64)
    order = None
    if not presorted and len(ts) > 1:                                   # candle core.py:60
        key = np.where(ts == F.NAT, I64_MAX, ts)
        fid = np.repeat(np.arange(nf, dtype=np.int64), lens)
        bad = (key[1:] < key[:-1]) & (fid[1:] == fid[:-1])
        if bad.any():
            order = np.lexsort((key, fid))
    pick = (lambda a: np.ascontiguousarray(a)) if order is None else (lambda a: np.ascontiguousarray(a[order]))
    t = lambda a: _to_device(a, dev)
    res = launch(t(off), t(pick(ts)), *(t(pick(big[c].to_numpy())) for c in TASK2_NUM),
                 bucket_ticks=n * 60 * F.ticks_per_second(unit), min_count=n, check_len=check_len, sym_valid=None,
                 want_index=True, trim=True)
    status, flags, out_off, goff, o_ts, o_sym, o_grp, oo, oh, ol, oc, ov = d2h_many(
        [res.sym_status, res.sym_flags, res.out_off, res.sym_group_off, res.ts, res.sym_row, res.group, res.open, res.high,
         res.low, res.close, res.volume])
    sym_row = o_sym.astype(np.int64)
    if order is not None:
        sym_row = np.where(sym_row >= 0, order[np.maximum(sym_row, 0)], -1)
    contract = np.repeat(np.arange(nf, dtype=np.int64), np.diff(out_off))
    data = {"symbol": _take_fill(big["symbol"].array, sym_row),
            "timestamp": o_ts.view(f"datetime64[{unit}]"),
            "open": oo, "high": oh, "low": ol, "close": oc, "volume": ov}
    out = pd.DataFrame(data, index=pd.Index(o_grp.astype(np.int64) - goff[contract], copy=False), copy=False)
    out["frequency"] = target_frequency                                 # candle core.py:92-94
    out["source_candles"] = n
    out["created_at"] = created_at
    out = out[list(F.CANDLE_OUT_ORDER)]
    results: List[Optional[pd.DataFrame]] = [None] * nf
    if whole_out is not None:                                           # the frame the results are slices of (see FrameSlices)
        whole_out.update(frame=out, off=np.asarray(out_off, dtype=np.int64), ok=np.asarray(status) == SYM_OK)
    if slices:
        for g in range(nf):
            if status[g] == SYM_OK:
                results[g] = out.iloc[int(out_off[g]):int(out_off[g + 1])]
    return results, [int(x) for x in flags]
