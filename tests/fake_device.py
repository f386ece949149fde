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


def bridge_plan(in_off, base, volume):
    """Stand-in of ``device.bridge_plan``: emitted candles / exponential draws before every symbol (no segment tables)."""
    from oracle import bridge_oracle as BO
    off = in_off.numpy()
    ok = BO.valid_rows(base.numpy())
    need = ok & BO.needs_volume(volume.numpy())
    c = np.concatenate([[0], np.cumsum(ok)]).astype(np.int64)
    ce = np.concatenate([[0], np.cumsum(need)]).astype(np.int64)
    return D.BridgePlan(torch.from_numpy(c[off]), torch.from_numpy(ce[off]), None, None, int(ok.sum()), int(need.sum()))


def bridge_ohlcv(in_off, base, volume, plan, strategy, *, u=None, e=None, z=None, base_spread_percent=0.002,
                 volatility_factor=1.5, **_):
    """Stand-in of ``device.bridge_ohlcv``: the reference-pinned row loop of oracle/bridge_oracle.py per symbol, each symbol
    reading its own stretch of the flat tapes (k uniforms per emitted row, one exponential per drawn volume)."""
    from iv_interpolation_b200 import _lib
    from oracle import bridge_oracle as BO
    off, out_off, e_off = in_off.numpy(), plan.out_off.numpy(), plan.e_off.numpy()
    k = _lib.BRIDGE_U_PER_ROW[_lib.BRIDGE_STRATEGY[strategy]]
    un = u.numpy() if u is not None else np.zeros(0)
    en = e.numpy() if e is not None else np.zeros(0)
    zn = z.numpy() if z is not None else np.zeros(0)
    names = ("open", "high", "low", "close", "volume", "source_price")
    parts, q = [], np.zeros(len(off) - 1, np.int32)
    for g in range(len(off) - 1):
        a, b = int(off[g]), int(off[g + 1])
        r = BO.bridge_ohlcv(strategy, base.numpy()[a:b], volume.numpy()[a:b], un[k * out_off[g]:k * out_off[g + 1]],
                            en[e_off[g]:e_off[g + 1]], zn[out_off[g]:out_off[g + 1]] if len(zn) else None,
                            base_spread_percent=base_spread_percent, volatility_factor=volatility_factor)
        assert len(r["row"]) == out_off[g + 1] - out_off[g]
        r["row"] = r["row"] + a
        q[g] = r["quality"]
        parts.append(r)
    cat = lambda key, dt: torch.from_numpy(np.concatenate([r[key] for r in parts]).astype(dt) if parts else np.zeros(0, dt))
    return D.BridgeBatch(plan.out_off, cat("row", np.int64), *(cat(n, np.float64) for n in names), torch.from_numpy(q), plan.total)
