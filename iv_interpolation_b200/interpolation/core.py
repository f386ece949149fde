"""IVInterpolator — same class, constructor and method as reference src/interpolation/core.py:9-85,
computed by the sm_100a kernels of libivx.so (kernel A).  No CPU fallback.

Error contract (reference core.py:26-28,36-39,49-51,76-78,83-85): never raises; insufficient or
invalid input, or any internal failure, logs and returns ``None``.
"""
from __future__ import annotations

import logging
from typing import List, Optional, Sequence

import numpy as np
import pandas as pd

from .. import frames as F

logger = logging.getLogger(__name__)

# Series.interpolate methods that reduce to np.interp on the positional index of the merged
# frame (reference core.py:61 passes `method` through verbatim).  'time' needs a DatetimeIndex the
# merged frame does not have -> the reference returns None; scipy-backed methods are not offered.
_LINEAR_METHODS = ("linear", "index", "values")


class IVInterpolator:
    """Core interpolation engine for IV data (B200 implementation)."""

    def __init__(self, method: str = "linear", min_points: int = 10):
        self.method = method
        self.min_points = min_points

    # -- single contract: the reference's entry point -------------------------------------------
    def interpolate_symbol(self, symbol_data: pd.DataFrame) -> Optional[pd.DataFrame]:
        if len(symbol_data) < self.min_points:                                   # core.py:26-28
            logger.warning(f"Insufficient data points: {len(symbol_data)} < {self.min_points}")
            return None
        try:
            out = self.interpolate_batch([symbol_data])[0]
            return out
        except Exception as e:                                                   # core.py:83-85
            logger.error(f"Interpolation failed: {e}")
            return None

    # -- many contracts in one launch --------------------------------------------------------------
    def interpolate_batch(self, frames: Sequence[pd.DataFrame], timing: Optional[dict] = None,
                          extra: Optional[dict] = None, whole_out: Optional[dict] = None) -> List[Optional[pd.DataFrame]]:
        """``[interpolate_symbol(f) for f in frames]`` with one plan + one exec launch for all.  ``whole_out`` (a dict)
        receives the one frame the results are slices of, when there is one (hostpath.FrameSlices): ``frame``, ``off``, ``ok``
        and ``slots`` (the positions in ``frames`` the offsets stand for)."""
        import torch
        from .. import device as D
        from .._lib import SYM_OK, IvxError, lib

        lib()                                          # fail loudly before touching the data
        if not torch.cuda.is_available():
            raise IvxError("no CUDA device: iv_interpolation_b200 has no CPU path")
        results: List[Optional[pd.DataFrame]] = [None] * len(frames)
        if self.method not in _LINEAR_METHODS:
            logger.error(f"Interpolation failed: method {self.method!r} is not available "
                         f"(supported: {_LINEAR_METHODS})")
            return results
        # frames in the reference's plain schema: one concat, one launch, one result frame (hostpath.py)
        from .. import hostpath as H
        big_enough = [i for i, df in enumerate(frames) if len(df) >= self.min_points]
        if len(big_enough) > 1 and H.task1_uniform([frames[i] for i in big_enough]):
            outs = H.interpolate_uniform([frames[i] for i in big_enough], self.min_points, timing=timing, extra=extra,
                                         whole_out=whole_out)
            if whole_out is not None:
                whole_out["slots"] = big_enough
            for i, out in zip(big_enough, outs):
                if out is None:
                    logger.debug("No valid data after interpolation or guard failed")
                results[i] = out
            return results
        inputs, srcs, slots = [], [], []
        for i, df in enumerate(frames):
            if len(df) < self.min_points:
                continue
            try:
                src = df.reset_index(drop=True)
                inp = F.interp_frame_to_arrays(src)
            except Exception as e:                     # missing column, bad dates, object dtype ...
                logger.error(f"Interpolation failed: {e}")
                continue
            # rows travel in time order (core.py:32); NaT last
            key = np.where(inp.date == F.NAT, np.iinfo(np.int64).max, inp.date)
            if len(key) > 1 and np.any(key[1:] < key[:-1]):
                order = np.argsort(key, kind="stable")
            else:
                order = None
            inputs.append((inp, order)); srcs.append(src); slots.append(i)
        if not inputs:
            return results

        # one launch per group of frames that share a column layout and a time unit
        groups = {}
        for j, (inp, order) in enumerate(inputs):
            sig = (inp.unit, tuple(inp.fill_num_names), tuple(inp.fill_idx_names))
            groups.setdefault(sig, []).append(j)
        dev = torch.device("cuda")
        for sig, members in groups.items():
            unit, num_names, idx_names = sig
            idx_names = list(idx_names)
            sym_first = ["symbol"] + [c for c in idx_names if c != "symbol"] if "symbol" in idx_names else None
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

            date = torch.from_numpy(cat(lambda p: p.date)).to(dev)
            icols = [torch.from_numpy(cat(lambda p, k=k: p.interp[k])).to(dev) for k in range(3)]
            fcols = [torch.from_numpy(cat(lambda p, k=k: p.fill_num[k])).to(dev) for k in range(len(num_names))]
            if sym_first is None:
                # `symbol` is a float column: its validity still drives dropna (core.py:74)
                k = list(num_names).index("symbol")
                masks = [cat(lambda p: ~np.isnan(p.fill_num[k]))] + \
                        [cat(lambda p, c=c: p.fill_idx_valid[p.fill_idx_names.index(c)]) for c in idx_names]
                order_names = ["symbol"] + idx_names
            else:
                masks = [cat(lambda p, c=c: p.fill_idx_valid[p.fill_idx_names.index(c)]) for c in sym_first]
                order_names = sym_first
            if len(fcols) > 8 or len(masks) > 4:        # only these frames fail (the reference would handle them: logged)
                logger.error(f"Interpolation failed: {len(fcols)} numeric and {len(masks)} non-numeric forward-fill columns exceed "
                             "one launch (8 / 4); convert string columns you do not need or split the frame")
                continue
            valid = [None if m.all() else torch.from_numpy(m.astype(np.uint8)).to(dev) for m in masks]
            first = inputs[members[0]][0]
            res = D.interp_linear_ffill(
                torch.from_numpy(off).to(dev), date, icols, fcols, valid,
                step=first.step, max_span=first.max_span, min_points=self.min_points,
                want_src_row=True, want_pos=True, want_keep=True, want_is_interpolated=False)
            status = res.sym_status.cpu().numpy()
            out_off = res.out_off.cpu().numpy()
            o_ts = res.ts.cpu().numpy()
            o_interp = [t.cpu().numpy() for t in res.interp]
            o_fill = [t.cpu().numpy() for t in res.fill]
            o_idx = [t.cpu().numpy() for t in res.idx]
            o_src = res.src_row.cpu().numpy()
            o_pos = res.pos.cpu().numpy()
            o_keep = res.keep.cpu().numpy().astype(bool)
            for g, j in enumerate(members):
                if status[g] != SYM_OK:
                    continue
                inp, order = inputs[j]
                a, b = int(out_off[g]), int(out_off[g + 1])
                keep = o_keep[a:b]
                if not keep.any():                                              # core.py:76-78
                    logger.warning("No valid data after interpolation")
                    continue
                sl = (lambda x: x[a:b]) if keep.all() else (lambda x: x[a:b][keep])
                base = int(off[g])

                def to_orig(ix):
                    """global sorted-row index -> row of the caller's frame (-1 stays -1)."""
                    loc = np.where(ix >= 0, ix - base, 0)
                    if order is not None:
                        loc = order[loc]
                    return np.where(ix >= 0, loc, -1)

                idx_by_name = {c: to_orig(sl(o_idx[k])) for k, c in enumerate(order_names)}
                results[slots[j]] = F.interp_arrays_to_frame(
                    srcs[j], inp, sl(o_ts), [sl(c) for c in o_interp], [sl(c) for c in o_fill],
                    [idx_by_name[c] for c in inp.fill_idx_names], to_orig(sl(o_src)), sl(o_pos).astype(np.int64))
                logger.debug(f"Interpolated {inp.n} → {int(keep.sum())} rows")
        if extra:                                      # the general path: per frame
            for r in results:
                if r is not None:
                    for name, value in extra.items():
                        r[name] = value
        return results
