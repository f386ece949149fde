"""Oracle-backed stand-ins for the two device launches the host paths make.  TEST INFRASTRUCTURE ONLY.

The batched host path (iv_interpolation_b200/hostpath.py) is pure data movement around one launch; on the CPU-only test tier the
launch is replaced by these functions, which run the reference-pinned numpy oracle per symbol and return the same result objects
the device wrappers return (CPU tensors).  The product never imports this module."""
import numpy as np
import torch

from iv_interpolation_b200 import device as D
from oracle import np_oracle as O


def interp_linear_ffill(src_off, ts, interp_cols, fill_cols=(), idx_valid=(), *, step, max_span, max_points=100_000, min_points=10,
                        want_src_row=False, want_pos=False, want_keep=False, want_is_interpolated=True):
    off = src_off.numpy()
    n_sym = len(off) - 1
    n = int(off[-1])
    masks = [np.ones(n, bool) if m is None else m.numpy().astype(bool) for m in idx_valid]
    parts, status, counts = [], np.zeros(n_sym, np.uint8), np.zeros(n_sym, np.int64)
    for g in range(n_sym):
        a, b = int(off[g]), int(off[g + 1])
        r = O.interp_linear_ffill(ts.numpy()[a:b], [c.numpy()[a:b] for c in interp_cols], [c.numpy()[a:b] for c in fill_cols],
                                  [m[a:b] for m in masks], step=step, max_span=max_span, max_points=max_points, min_points=min_points)
        status[g] = r["status"]
        if r["status"] == O.SYM_OK:
            counts[g] = len(r["date"])
            r["base"] = a
            parts.append(r)
    out_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    cat = lambda f, dt: torch.from_numpy(np.ascontiguousarray(np.concatenate([f(r) for r in parts]) if parts else np.zeros(0)).astype(dt))
    glob = lambda r, ix: np.where(ix >= 0, ix + r["base"], -1)
    return D.InterpBatch(
        torch.from_numpy(out_off), torch.from_numpy(status), torch.zeros(n_sym, dtype=torch.int32), cat(lambda r: r["date"], np.int64),
        [cat(lambda r, k=k: r["interp"][k], np.float64) for k in range(3)],
        [cat(lambda r, k=k: r["fill"][k], np.float64) for k in range(len(fill_cols))],
        [cat(lambda r, k=k: glob(r, r["str_src"][k]), np.int32) for k in range(len(idx_valid))],
        cat(lambda r: glob(r, r["src_row"]), np.int32) if want_src_row else None,
        cat(lambda r: r["index"], np.int32) if want_pos else None,
        torch.ones(int(out_off[-1]), dtype=torch.uint8) if want_keep else None,
        torch.zeros(int(out_off[-1]), dtype=torch.uint8) if want_is_interpolated else None, int(out_off[-1]))


def ohlcv_reduce(in_off, ts, o, h, l, c, v, *, bucket_ticks, min_count, check_len=True, sym_valid=None, want_index=False,
                 capacity=None, trim=True):
    off = in_off.numpy()
    n_sym = len(off) - 1
    parts, status, counts, flags = [], np.zeros(n_sym, np.uint8), np.zeros(n_sym, np.int64), np.zeros(n_sym, np.int32)
    for g in range(n_sym):
        a, b = int(off[g]), int(off[g + 1])
        cols = [x.numpy()[a:b] for x in (o, h, l, c, v)]
        r = O.ohlcv_reduce(ts.numpy()[a:b], *cols, bucket_ticks=bucket_ticks, min_count=min_count, presorted=True, check_len=check_len,
                           sym_valid=None if sym_valid is None else sym_valid.numpy()[a:b].astype(bool))
        status[g] = r["status"]
        if r["status"] == O.SYM_OK:
            counts[g] = len(r["ts"])
            r["base"] = a
            parts.append(r)
            fl = O.validate_flags(*cols)
            if len(r["ts"]):
                fl |= O.validate_flags(*(r[k] for k in ("open", "high", "low", "close", "volume"))) << 3
            flags[g] = fl
    out_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    cat = lambda f, dt: torch.from_numpy(np.ascontiguousarray(np.concatenate([f(r) for r in parts]) if parts else np.zeros(0)).astype(dt))
    res = D.CandleBatch(torch.from_numpy(out_off), cat(lambda r: r["ts"], np.int64), *(cat(lambda r, k=k: r[k], np.float64)
                        for k in ("open", "high", "low", "close", "volume")), cat(lambda r: r["count"], np.int32),
                        torch.from_numpy(status), torch.from_numpy(flags),
                        cat(lambda r: np.where(r["sym_row"] >= 0, r["sym_row"] + r["base"], -1), np.int64),
                        cat(lambda r: r["group_index"], np.int64), torch.zeros(n_sym + 1, dtype=torch.int64), int(out_off[-1]))
    return res
