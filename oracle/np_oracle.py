"""numpy restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function cites the reference lines (relative to /root/reference) it restates.
The arithmetic of kernels A and B lives in two un-vendored third-party packages,
floor-pinned only by the reference (requirements.txt:2-3: pandas>=1.5.0,
numpy>=1.24.0); the build container has pandas 3.0.2 / numpy 2.3.5 and those are the
versions the golden vectors under tests/golden/ were produced with:

  * ``Series.interpolate('linear')`` -> pandas.core.missing._interpolate_1d ->
    ``np.interp`` on the POSITIONAL index, leading NaNs preserved, trailing NaNs
    flat-filled (limit_direction='forward');
  * ``np.interp`` -> numpy/_core/src/multiarray/compiled_base.c ``arr_interp``:
    ``slope = (y1-y0)/(x1-x0); y = slope*(x-x0) + y0`` with separate multiply and
    add (no FMA in numpy's x86-64 baseline build) and a two-step NaN fallback;
  * ``groupby.agg`` -> pandas/_libs/groupby.pyx ``group_sum`` (Kahan-compensated,
    NaN skipped, compensation reset to 0 when it becomes NaN), ``group_nth``-style
    first/last non-null, ``group_max/min`` NaN-skipping.

All timestamps are int64 ticks in the caller's datetime64 unit; NaT is INT64_MIN.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np

NAT = np.iinfo(np.int64).min
I64_MAX = np.iinfo(np.int64).max

# reference src/interpolation/core.py:58 and :64-65
INTERP_COLS = ("iv", "underlying_price", "time_to_maturity")
FILL_NUM_COLS = ("strike", "interest_rate", "mark_price", "index_price", "volume", "quote_volume")
FILL_STR_COLS = ("symbol", "callput", "record_time")

# status codes shared with include/ivx.h (IVX_SYM_*)
SYM_OK = 0
SYM_TOO_FEW = 1        # core.py:26-28 / candle core.py:64-66
SYM_SPAN = 2           # core.py:36-39
SYM_TOO_LONG = 3       # core.py:49-51
SYM_EMPTY_OUT = 4      # core.py:76-78
SYM_NO_DATES = 5       # all dates NaT -> date_range raises -> core.py:83-85


# --------------------------------------------------------------------------------------
# np.interp, restated (numpy compiled_base.c arr_interp)
# --------------------------------------------------------------------------------------
def interp_restated(x: np.ndarray, xp: np.ndarray, fp: np.ndarray) -> np.ndarray:
    """``np.interp(x, xp, fp)`` spelled out so the C and CUDA versions can follow it.

    xp strictly increasing, no NaN in xp/fp (pandas passes only valid knots).
    """
    x = np.asarray(x, dtype=np.float64)
    xp = np.asarray(xp, dtype=np.float64)
    fp = np.asarray(fp, dtype=np.float64)
    n = len(xp)
    if n == 1:
        return np.full(x.shape, fp[0])
    j = np.searchsorted(xp, x, side="right") - 1          # largest j with xp[j] <= x
    jj = np.clip(j, 0, n - 2)
    with np.errstate(all="ignore"):
        slope = (fp[jj + 1] - fp[jj]) / (xp[jj + 1] - xp[jj])
        r = slope * (x - xp[jj]) + fp[jj]
        bad = np.isnan(r)
        r2 = slope * (x - xp[jj + 1]) + fp[jj + 1]
        r = np.where(bad, r2, r)
        r = np.where(bad & np.isnan(r) & (fp[jj] == fp[jj + 1]), fp[jj], r)
    r = np.where(x == xp[jj], fp[jj], r)
    r = np.where(j >= n - 1, fp[n - 1], r)                # x >= last knot -> right value
    r = np.where(j < 0, fp[0], r)                         # x <  first knot -> left value
    return r


def _pandas_linear_interpolate(y: np.ndarray) -> np.ndarray:
    """``Series.interpolate('linear')`` on a default RangeIndex (core.py:58-61)."""
    y = y.copy()
    valid = ~np.isnan(y)
    if not valid.any() or valid.all():
        return y
    pos = np.arange(len(y), dtype=np.float64)
    inv = ~valid
    y[inv] = interp_restated(pos[inv], pos[valid], y[valid])
    first = int(np.argmax(valid))
    y[:first] = np.nan                                     # leading NaNs are preserved
    return y


def _ffill_index(valid: np.ndarray) -> np.ndarray:
    """index of the last valid position <= p, or -1 (``fillna(method='ffill')``, core.py:66-68)."""
    idx = np.where(valid, np.arange(len(valid), dtype=np.int64), -1)
    return np.maximum.accumulate(idx) if len(idx) else idx


def stable_time_order(ts: np.ndarray) -> np.ndarray:
    """``sort_values`` order with NaT last; ties keep input order.

    The reference's sort (core.py:32, candle core.py:60) is numpy's unstable
    quicksort, so tie order is undefined there; parity tapes have unique stamps.
    """
    key = np.where(ts == NAT, I64_MAX, ts)
    return np.argsort(key, kind="stable")


# --------------------------------------------------------------------------------------
# Kernel A: IVInterpolator.interpolate_symbol  (src/interpolation/core.py:16-85)
# --------------------------------------------------------------------------------------
def interp_linear_ffill(
    date: np.ndarray,
    interp_cols: Sequence[np.ndarray],
    fill_cols: Sequence[np.ndarray],
    str_valid: Sequence[np.ndarray],
    *,
    step: int,
    max_span: int,
    max_points: int = 100_000,
    min_points: int = 10,
) -> Dict[str, object]:
    """One contract.  Returns ``{'status': code}`` and, when status==0, the output columns.

    date        int64[n] ticks (input order, NaT allowed)
    interp_cols 3 float64[n]  (iv, underlying_price, time_to_maturity)  core.py:58
    fill_cols   float64[n] each (numeric forward-fill columns)           core.py:64-65
    str_valid   bool[n] each: non-null mask of a string column; FIRST one is ``symbol``
    step        60 s in ticks (core.py:45); max_span 30 d in ticks (core.py:37)

    Output arrays (length m = rows surviving dropna, core.py:74):
      date int64, interp (3,m), fill (k,m), str_src (s,m) = ORIGINAL input-row index whose
      string is carried (or -1), src_row (m,) original row at that grid point or -1.
    """
    date = np.asarray(date, dtype=np.int64)
    n = len(date)
    if n < min_points:                                        # core.py:26-28
        return {"status": SYM_TOO_FEW}
    order = stable_time_order(date)                           # core.py:32
    d = date[order]
    ok_t = d != NAT
    if not ok_t.any():
        return {"status": SYM_NO_DATES}
    t_min = int(d[ok_t][0])
    t_max = int(d[ok_t][-1])
    if t_max - t_min > max_span:                              # core.py:36-39 (strict >)
        return {"status": SYM_SPAN}
    n_grid = (t_max - t_min) // step + 1                      # core.py:42-46
    if n_grid > max_points:                                   # core.py:49-51
        return {"status": SYM_TOO_LONG}

    # left join timeline <- rows (core.py:54-55): a row lands iff it is exactly on the grid;
    # m rows on one stamp produce m consecutive output positions.
    off = d - t_min
    on = ok_t & (off % step == 0)
    rows = np.flatnonzero(on)                                 # sorted-row indices on the grid
    k = (off[rows] // step).astype(np.int64)
    dup = np.r_[False, k[1:] == k[:-1]] if len(k) else np.zeros(0, bool)
    pos = k + np.cumsum(dup)
    n_pos = int(n_grid + dup.sum())
    is_dup_pos = np.zeros(n_pos, bool)
    is_dup_pos[pos[dup]] = True
    grid_idx = np.arange(n_pos, dtype=np.int64) - np.cumsum(is_dup_pos)
    out_date = t_min + grid_idx * step
    src_sorted = np.full(n_pos, -1, dtype=np.int64)
    src_sorted[pos] = rows

    def spread(col: np.ndarray) -> np.ndarray:
        y = np.full(n_pos, np.nan)
        y[pos] = np.asarray(col, dtype=np.float64)[order][rows]
        return y

    interp_out = [_pandas_linear_interpolate(spread(c)) for c in interp_cols]   # core.py:58-61
    fill_out = []
    for c in fill_cols:                                                        # core.py:64-68
        y = spread(c)
        idx = _ffill_index(~np.isnan(y))
        fill_out.append(np.where(idx >= 0, y[np.maximum(idx, 0)], np.nan))
    str_src = []
    for v in str_valid:
        ok = np.zeros(n_pos, bool)
        ok[pos] = np.asarray(v, dtype=bool)[order][rows]
        idx = _ffill_index(ok)
        str_src.append(np.where(idx >= 0, order[src_sorted[np.maximum(idx, 0)]], -1))

    keep = np.ones(n_pos, bool)                                                # core.py:74
    if str_src:
        keep &= str_src[0] >= 0
    for y in interp_out:
        keep &= ~np.isnan(y)
    if not keep.any():                                                         # core.py:76-78
        return {"status": SYM_EMPTY_OUT}
    return {
        "status": SYM_OK,
        "date": out_date[keep],
        "interp": np.stack([y[keep] for y in interp_out]) if interp_out else np.zeros((0, int(keep.sum()))),
        "fill": np.stack([y[keep] for y in fill_out]) if fill_out else np.zeros((0, int(keep.sum()))),
        "str_src": np.stack([s[keep] for s in str_src]).astype(np.int64) if str_src else np.zeros((0, int(keep.sum())), np.int64),
        "src_row": np.where(src_sorted >= 0, order[np.maximum(src_sorted, 0)], -1)[keep],
        "n_pos": n_pos,
        "index": np.flatnonzero(keep),                        # positional labels dropna leaves
        "keep_is_suffix": bool(keep[int(np.argmax(keep)):].all()),
    }


# --------------------------------------------------------------------------------------
# Kernel B: CandleReconstructor.reconstruct_symbol_candles (src/candle_reconstruction/core.py:36-106)
# --------------------------------------------------------------------------------------
def kahan_group_sum(values: np.ndarray, starts: np.ndarray, ends: np.ndarray) -> np.ndarray:
    """pandas ``group_sum`` (Kahan) per [start,end) run, NaN skipped, empty -> 0.0 (core.py:78)."""
    nb = len(starts)
    s = np.zeros(nb)
    comp = np.zeros(nb)
    lens = ends - starts
    with np.errstate(all="ignore"):
        for kk in range(int(lens.max()) if nb else 0):
            act = np.flatnonzero(lens > kk)
            val = values[starts[act] + kk]
            ok = ~np.isnan(val)
            a = act[ok]
            y = val[ok] - comp[a]
            t = s[a] + y
            c = t - s[a] - y
            c = np.where(np.isnan(c), 0.0, c)             # GH#53606: inf keeps the sum at inf
            comp[a] = c
            s[a] = t
    return s


def ohlcv_reduce(
    ts: np.ndarray,
    o: np.ndarray,
    h: np.ndarray,
    l: np.ndarray,
    c: np.ndarray,
    v: np.ndarray,
    *,
    bucket_ticks: int,
    min_count: int,
    sym_valid: Optional[np.ndarray] = None,
    presorted: bool = False,
    check_len: bool = True,
) -> Dict[str, object]:
    """One symbol.  Follows candle core.py:59-89 (and complete_pipeline.py:685-710 when
    ``presorted=True, check_len=False``).

    Returns status and, for kept buckets: ts (bucket start), open, high, low, close, volume,
    count (raw rows incl. NaN rows, core.py:87), sym_row (original row index of the first
    non-null symbol of the bucket, -1 if none; core.py:79), first_row (original index of the
    bucket's first row).
    """
    ts = np.asarray(ts, dtype=np.int64)
    n = len(ts)
    if n == 0:                                                # core.py:47-49
        return {"status": SYM_TOO_FEW}
    if check_len and n < min_count:                           # core.py:64-66
        return {"status": SYM_TOO_FEW}
    order = np.arange(n) if presorted else stable_time_order(ts)   # core.py:60
    good = ts[order] != NAT                                   # groupby drops NaT keys
    order = order[good]
    t = ts[order]
    cols = [np.asarray(a, dtype=np.float64)[order] for a in (o, h, l, c, v)]
    sv = np.ones(len(order), bool) if sym_valid is None else np.asarray(sym_valid, bool)[order]
    b = np.floor_divide(t, bucket_ticks)                      # core.py:70 (epoch-anchored floor)
    if presorted and len(b):
        # groupby(sort=True) on an unsorted frame would regroup; the no-sort entry is only
        # defined (and only used, complete_pipeline.py:645-648 ORDER BY) on sorted input.
        assert np.all(b[1:] >= b[:-1]), "presorted=True requires non-decreasing timestamps"
    m = len(t)
    heads = np.r_[True, b[1:] != b[:-1]] if m else np.zeros(0, bool)
    starts = np.flatnonzero(heads)
    ends = np.r_[starts[1:], m] if m else starts
    count = ends - starts
    ar = np.arange(m, dtype=np.int64)

    def first_valid(mask):
        idx = np.where(mask, ar, m)
        return np.minimum.reduceat(idx, starts) if len(starts) else idx[:0]

    def last_valid(mask):
        idx = np.where(mask, ar, -1)
        return np.maximum.reduceat(idx, starts) if len(starts) else idx[:0]

    O, H, L, C, V = cols
    nb = len(starts)
    if nb:
        fo = first_valid(~np.isnan(O))
        lc = last_valid(~np.isnan(C))
        fs = first_valid(sv)
        open_ = np.where(fo < m, O[np.minimum(fo, m - 1)], np.nan)      # core.py:74 'first'
        close_ = np.where(lc >= 0, C[np.maximum(lc, 0)], np.nan)         # core.py:77 'last'
        with np.errstate(all="ignore"):
            high_ = np.fmax.reduceat(H, starts)                           # core.py:75 'max'
            low_ = np.fmin.reduceat(L, starts)                            # core.py:76 'min'
        vol_ = kahan_group_sum(V, starts, ends)                           # core.py:78 'sum'
        sym_row = np.where(fs < m, order[np.minimum(fs, m - 1)], -1)
    else:
        open_ = close_ = high_ = low_ = vol_ = np.zeros(0)
        sym_row = np.zeros(0, np.int64)
    keep = count >= min_count                                 # core.py:87-89
    return {
        "status": SYM_OK,
        "ts": (b[starts] * bucket_ticks)[keep] if nb else np.zeros(0, np.int64),
        "open": open_[keep], "high": high_[keep], "low": low_[keep],
        "close": close_[keep], "volume": vol_[keep],
        "count": count[keep].astype(np.int64),
        "sym_row": sym_row[keep].astype(np.int64),
        "first_row": order[starts][keep].astype(np.int64) if nb else np.zeros(0, np.int64),
        "group_index": np.flatnonzero(keep),                  # the (not reset) index pandas returns
    }


def validate_candles(o, h, l, c, v) -> bool:
    """CandleReconstructor.validate_candle_data (candle core.py:108-144)."""
    o, h, l, c, v = (np.asarray(a, dtype=np.float64) for a in (o, h, l, c, v))
    if len(o) == 0:                                           # :110-111
        return False
    with np.errstate(invalid="ignore"):
        bad = (h < l) | (h < o) | (h < c) | (l > o) | (l > c)  # :115-119 (NaN compares False)
        if bad.any():
            return False
        if (v < 0).any():                                     # :127-130
            return False
    for a in (o, h, l, c):                                    # :133-137
        if np.isnan(a).any():
            return False
    return True


def validate_flags(o, h, l, c, v) -> int:
    """Bit flags the CUDA kernel reports per symbol (include/ivx.h IVX_FLAG_*)."""
    o, h, l, c, v = (np.asarray(a, dtype=np.float64) for a in (o, h, l, c, v))
    f = 0
    with np.errstate(invalid="ignore"):
        if ((h < l) | (h < o) | (h < c) | (l > o) | (l > c)).any():
            f |= 1
        if (v < 0).any():
            f |= 2
    if any(np.isnan(a).any() for a in (o, h, l, c)):
        f |= 4
    return f


# --------------------------------------------------------------------------------------
# Black-Scholes conventions (src/interpolation/greeks.py:12-43) and the IV inversion oracle
# --------------------------------------------------------------------------------------
def _ndtr(x):
    from scipy.special import ndtr
    return ndtr(x)


def bs_greeks(S, K, T, r, sigma, is_call=True, reference_put_rho_sign=True):
    """greeks.py:21-35.  ``reference_put_rho_sign`` keeps the reference's +sign on put rho (:35)."""
    S, K, T, r, sigma = (np.asarray(a, dtype=np.float64) for a in (S, K, T, r, sigma))
    sq = np.sqrt(T)
    d1 = (np.log(S / K) + (r + 0.5 * sigma ** 2) * T) / (sigma * sq)      # :21
    d2 = d1 - sigma * sq                                                   # :22
    pdf = np.exp(-0.5 * d1 * d1) / math.sqrt(2.0 * math.pi)
    disc = np.exp(-r * T)
    call = np.asarray(is_call, dtype=bool)
    delta = np.where(call, _ndtr(d1), _ndtr(d1) - 1.0)                     # :25,:29
    theta = np.where(
        call,
        (-S * pdf * sigma / (2 * sq) - r * K * disc * _ndtr(d2)) / 365,     # :26-27
        (-S * pdf * sigma / (2 * sq) + r * K * disc * _ndtr(-d2)) / 365,    # :30-31
    )
    gamma = pdf / (S * sigma * sq)                                         # :33
    vega = S * pdf * sq / 100                                              # :34
    rho = K * T * disc * np.where(call, _ndtr(d2), _ndtr(-d2)) / 100       # :35 (put sign as shipped)
    if not reference_put_rho_sign:
        rho = np.where(call, rho, -rho)
    return {"delta": delta, "gamma": gamma, "theta": theta, "vega": vega, "rho": rho}


def bs_price(S, K, T, r, sigma, is_call):
    """Black-Scholes price under greeks.py:21-22 conventions (continuous r, no dividend, T in years)."""
    S, K, T, r, sigma = (np.asarray(a, dtype=np.float64) for a in (S, K, T, r, sigma))
    sq = np.sqrt(T)
    d1 = (np.log(S / K) + (r + 0.5 * sigma * sigma) * T) / (sigma * sq)
    d2 = d1 - sigma * sq
    disc = np.exp(-r * T)
    callp = S * _ndtr(d1) - K * disc * _ndtr(d2)
    putp = K * disc * _ndtr(-d2) - S * _ndtr(-d1)
    return np.where(np.asarray(is_call, dtype=bool), callp, putp)


IV_SIGMA_SQRT_T_MAX = 40.0     # fp64 cannot represent a price strictly below the upper bound past ~17


def implied_vol_domain(price, S, K, T, r, is_call):
    """The agreed invertible domain: finite positive inputs and
    lower < price < upper in fp64, with lower/upper the no-arbitrage bounds."""
    price, S, K, T, r = (np.asarray(a, dtype=np.float64) for a in (price, S, K, T, r))
    call = np.asarray(is_call, dtype=bool)
    with np.errstate(all="ignore"):
        kd = K * np.exp(-r * T)
        intrinsic = np.where(call, S - kd, kd - S)
        lower = np.maximum(intrinsic, 0.0)
        upper = np.where(call, S, kd)
        ok = np.isfinite(price) & np.isfinite(S) & np.isfinite(K) & np.isfinite(T) & np.isfinite(r)
        ok &= (S > 0) & (K > 0) & (T > 0) & (price > lower) & (price < upper)
    return ok


def implied_vol(price, S, K, T, r, is_call, xtol=1e-15):
    """scipy.optimize.brentq inversion, one quote at a time (correctness oracle, not the reference:
    the reference contains no IV solver).

    The root is bracketed on the out-of-the-money side: an in-the-money quote is first moved
    across put-call parity, ``otm = price - (S - K e^{-rT})`` (resp. ``+``), which is what makes
    deep in-the-money quotes solvable at all.  Returns NaN outside ``implied_vol_domain``.
    """
    from scipy.optimize import brentq

    price, S, K, T, r = (np.atleast_1d(np.asarray(a, dtype=np.float64)) for a in (price, S, K, T, r))
    call = np.atleast_1d(np.asarray(is_call, dtype=bool))
    price, S, K, T, r, call = np.broadcast_arrays(price, S, K, T, r, call)
    ok = implied_vol_domain(price, S, K, T, r, call)
    out = np.full(price.shape, np.nan)
    for i in np.flatnonzero(ok):
        Si, Ki, Ti, ri, pi, ci = float(S[i]), float(K[i]), float(T[i]), float(r[i]), float(price[i]), bool(call[i])
        kd = Ki * math.exp(-ri * Ti)
        intrinsic = (Si - kd) if ci else (kd - Si)
        if intrinsic > 0.0:            # in the money: solve the OTM twin
            target = pi - intrinsic
            ci = not ci
        else:
            target = pi
        if not target > 0.0:
            continue
        sq = math.sqrt(Ti)

        def f(sig):
            return float(bs_price(Si, Ki, Ti, ri, sig, ci)) - target

        hi = IV_SIGMA_SQRT_T_MAX / sq
        lo = 1e-12 / sq
        try:
            if f(lo) >= 0.0 or f(hi) <= 0.0:
                continue
            out[i] = brentq(f, lo, hi, xtol=xtol, rtol=8.881784197001252e-16, maxiter=500)
        except (ValueError, RuntimeError):
            continue
    return out


# --------------------------------------------------------------------------------------
# Natural cubic spline oracle (north_star kernel 2; no reference code)
# --------------------------------------------------------------------------------------
EXTRAP_NAN = 0      # NaN outside [x_first, x_last]
EXTRAP_POLY = 1     # continue the end polynomials (scipy CubicSpline default)
EXTRAP_FLAT = 2     # clamp to the end values


def natural_spline_eval(x, y, xq, extrap=EXTRAP_POLY):
    """scipy.interpolate.CubicSpline(bc_type='natural') on the knots whose x and y are finite.

    <2 valid knots -> NaN everywhere; 2 valid knots -> the straight line through them.
    """
    from scipy.interpolate import CubicSpline

    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xq = np.asarray(xq, dtype=np.float64)
    ok = np.isfinite(x) & np.isfinite(y)
    xv, yv = x[ok], y[ok]
    if len(xv) < 2:
        return np.full(xq.shape, np.nan)
    if len(xv) == 2:
        slope = (yv[1] - yv[0]) / (xv[1] - xv[0])
        out = yv[0] + slope * (xq - xv[0])
    else:
        out = CubicSpline(xv, yv, bc_type="natural", extrapolate=True)(xq)
    lo, hi = xv[0], xv[-1]
    if extrap == EXTRAP_NAN:
        out = np.where((xq < lo) | (xq > hi), np.nan, out)
    elif extrap == EXTRAP_FLAT:
        out = np.where(xq < lo, yv[0], np.where(xq > hi, yv[-1], out))
    return np.where(np.isnan(xq), np.nan, out)


def natural_spline_moments(x, y):
    """Second derivatives M of the natural spline by the Thomas algorithm (what the C oracle does)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(x)
    M = np.zeros(n)
    if n < 3:
        return M
    h = np.diff(x)
    sl = np.diff(y) / h
    cp = np.zeros(n)
    dp = np.zeros(n)
    for i in range(1, n - 1):
        a, b, c = h[i - 1], 2.0 * (h[i - 1] + h[i]), h[i]
        d = 6.0 * (sl[i] - sl[i - 1])
        den = b - a * cp[i - 1]
        cp[i] = c / den
        dp[i] = (d - a * dp[i - 1]) / den
    for i in range(n - 2, 0, -1):
        M[i] = dp[i] - cp[i] * M[i + 1]
    return M
