"""CPU oracle of the north_star surface pipeline.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

quotes -> implied vol -> natural cubic spline across strike -> across maturity -> OHLC per time bucket.
The reference contains none of this ("parity unpinned" by the reference): the
oracle is scipy (brentq, or the vectorised safeguarded Newton below which is held to brentq by
tests/test_oracle_surface.py; CubicSpline(bc_type='natural')) and, for the candles, the numpy
restatement of reference src/candle_reconstruction/core.py:68-89 in oracle/np_oracle.py.
"""
from __future__ import annotations

import numpy as np

from oracle import np_oracle as O


def implied_vol_fast(price, S, K, T, r, is_call, max_iter: int = 200):
    """Vectorised inversion on the out-of-the-money twin: Newton steps on the price, kept inside a
    shrinking bracket by bisection.  Same domain and NaN mask as ``np_oracle.implied_vol``."""
    price, S, K, T, r = (np.asarray(a, dtype=np.float64) for a in (price, S, K, T, r))
    call = np.asarray(is_call, dtype=bool)
    ok = O.implied_vol_domain(price, S, K, T, r, call)
    out = np.full(price.shape, np.nan)
    idx = np.flatnonzero(ok)
    if len(idx) == 0:
        return out
    p, s_, k_, t_, r_, c_ = price[idx], S[idx], K[idx], T[idx], r[idx], call[idx]
    kd = k_ * np.exp(-r_ * t_)
    intrinsic = np.where(c_, s_ - kd, kd - s_)
    itm = intrinsic > 0.0
    target = np.where(itm, p - intrinsic, p)
    c_ = np.where(itm, ~c_, c_)
    sq = np.sqrt(t_)
    lo = 1e-12 / sq
    hi = O.IV_SIGMA_SQRT_T_MAX / sq
    with np.errstate(all="ignore"):
        good = (target > 0.0) & (O.bs_price(s_, k_, t_, r_, lo, c_) < target) & (O.bs_price(s_, k_, t_, r_, hi, c_) > target)
        # sigma_c = the inflection point of price(sigma); below it the price is log-concave in the wings and
        # Newton on ln(price) converges fast, above it Newton on the price itself is monotone
        sig_c = np.maximum(np.sqrt(2.0 * np.abs(np.log(s_ / kd)) / t_), lo * 2)
        use_log = O.bs_price(s_, k_, t_, r_, sig_c, c_) > target
        ltarget = np.log(np.where(target > 0, target, 1.0))
        sig = np.clip(sig_c, lo, hi)
        active = good.copy()
        for _ in range(max_iter):
            if not active.any():
                break
            a = np.flatnonzero(active)
            pr = O.bs_price(s_[a], k_[a], t_[a], r_[a], sig[a], c_[a])
            d1 = (np.log(s_[a] / k_[a]) + (r_[a] + 0.5 * sig[a] ** 2) * t_[a]) / (sig[a] * sq[a])
            vega = s_[a] * np.exp(-0.5 * d1 * d1) * 0.3989422804014327 * sq[a]
            f = pr - target[a]
            step = np.where(use_log[a], (np.log(pr) - ltarget[a]) * pr / vega, f / vega)
            hi[a] = np.where(f > 0, sig[a], hi[a])
            lo[a] = np.where(f > 0, lo[a], sig[a])
            new = sig[a] - step
            bad = ~np.isfinite(new) | (new <= lo[a]) | (new >= hi[a])
            new = np.where(bad, 0.5 * (lo[a] + hi[a]), new)
            new = np.where(f == 0.0, sig[a], new)                     # landed on the root: keep it
            done = (np.abs(new - sig[a]) <= 4e-16 * new) | (hi[a] - lo[a] <= 4e-16 * hi[a]) | (f == 0.0)
            sig[a] = new
            active[a[done]] = False
    res = np.where(good, sig, np.nan)
    out[idx] = res
    return out


def smile_stage(K, iv, n_t, n_exp, n_strike, grid_k, extrap=O.EXTRAP_FLAT):
    """Natural cubic spline across strike per (timestamp, expiry) slice -> [n_t][n_exp][len(grid_k)]."""
    K3 = np.asarray(K).reshape(n_t, n_exp, n_strike)
    iv3 = np.asarray(iv).reshape(n_t, n_exp, n_strike)
    smile = np.empty((n_t, n_exp, len(grid_k)))
    for t in range(n_t):
        for e in range(n_exp):
            smile[t, e] = O.natural_spline_eval(K3[t, e], iv3[t, e], grid_k, extrap)
    return smile


def maturity_stage(T, smile, n_t, n_exp, n_strike, grid_m, extrap=O.EXTRAP_FLAT):
    """Natural cubic spline across maturity per (timestamp, grid strike); knots at T[t][e][0]
    -> [n_t][len(grid_m)][n_grid_k]."""
    T3 = np.asarray(T).reshape(n_t, n_exp, n_strike)
    nk = smile.shape[2]
    surface = np.empty((n_t, len(grid_m), nk))
    for t in range(n_t):
        xm = T3[t, :, 0]
        for j in range(nk):
            surface[t, :, j] = O.natural_spline_eval(xm, smile[t, :, j], grid_m, extrap)
    return surface


def candle_stage(ts, surface, bucket_ticks, min_count):
    """OHLC of every surface point over epoch-anchored buckets: the rules of reference
    src/candle_reconstruction/core.py:68-89 (floor bucket, first/last non-null, NaN-skipping max/min, a bucket
    is kept iff it holds >= min_count stamps) applied to value = open = high = low = close."""
    import warnings
    ts = np.asarray(ts, dtype=np.int64)
    n_t = len(ts)
    b = ts // bucket_ticks                      # epoch-anchored floor (candle core.py:70)
    uniq, start = np.unique(b, return_index=True)
    edges = np.r_[start, n_t]
    count = np.diff(edges).astype(np.int32)
    flat = np.asarray(surface).reshape(n_t, -1)
    o = np.full((len(uniq), flat.shape[1]), np.nan); h = o.copy(); l = o.copy(); c = o.copy()
    cols = np.arange(flat.shape[1])
    for i in range(len(uniq)):
        if count[i] < min_count:                # incomplete bucket (candle core.py:87-89)
            continue
        seg = flat[edges[i]:edges[i + 1]]
        valid = ~np.isnan(seg)
        anyv = valid.any(axis=0)
        first = np.argmax(valid, axis=0)
        last = seg.shape[0] - 1 - np.argmax(valid[::-1], axis=0)
        o[i] = np.where(anyv, seg[first, cols], np.nan)     # first non-null
        c[i] = np.where(anyv, seg[last, cols], np.nan)      # last non-null
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            h[i] = np.nanmax(seg, axis=0)                   # NaN-skipping
            l[i] = np.nanmin(seg, axis=0)
    return dict(bucket_ts=uniq * bucket_ticks, count=count, open=o, high=h, low=l, close=c)


def surface_pipeline(ts, price, S, K, T, r, is_call, n_exp, n_strike, grid_k, grid_m, bucket_ticks, min_count,
                     extrap=O.EXTRAP_FLAT, iv_solver=None):
    """The whole chain on numpy arrays laid out time-major (q = (t*n_exp + e)*n_strike + k).
    Returns dict(iv, smile[n_t][n_exp][n_grid_k], surface[n_t][n_grid_m][n_grid_k], bucket_ts, count,
    open/high/low/close[n_buckets][n_grid_m*n_grid_k])."""
    n_t = len(ts)
    iv = (iv_solver or implied_vol_fast)(price, S, K, T, r, is_call)
    smile = smile_stage(K, iv, n_t, n_exp, n_strike, grid_k, extrap)
    surface = maturity_stage(T, smile, n_t, n_exp, n_strike, grid_m, extrap)
    out = candle_stage(ts, surface, bucket_ticks, min_count)
    out.update(iv=iv, smile=smile, surface=surface)
    return out
