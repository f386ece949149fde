#!/usr/bin/env python
"""Known-answer vectors for the two kernels the reference has no code for, computed with mpmath at 50 digits.

    python tools/gen_kats_mpmath.py        # writes tests/golden/iv_kats.json and tests/golden/spline_kats.json (~3 minutes)

* implied volatility: quotes across the regimes of tests/test_gpu_parity.py::test_implied_vol_regimes (moneyness 1e-7 .. e^3,
  maturities an hour .. 30 years, vols 1 % .. 500 %, negative and large rates, calls and puts).  The price is the
  Black-Scholes value (conventions of reference src/interpolation/greeks.py:21-22) at 50 digits ROUNDED TO DOUBLE — what a
  kernel receives — and ``sigma`` is the 50-digit root for THAT double price, rounded to double; ``dsigma_per_rel`` (price / vega out of the money,
  max(price, S, K) / vega in the money) rides along so that the tests can state how well the double price determines sigma.
* natural cubic spline: ragged ladders (2 .. 333 knots, uneven spacing), second derivatives by Thomas' algorithm and values at
  query points inside and outside the ladder (polynomial continuation of the end segments), all at 50 digits.

These pin the ORACLES (scipy brentq / CubicSpline) and the kernels to something that does not move with the installed scipy.
Parity with the reference remains unpinned for both: the reference has no such code.
"""
from __future__ import annotations

import json
import os

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
mp.mp.dps = 50


def ncdf(z):
    return mp.erfc(-z / mp.sqrt(2)) / 2


def bs_price(S, K, T, r, sigma, call):
    sq = sigma * mp.sqrt(T)
    d1 = (mp.log(S / K) + (r + sigma * sigma / 2) * T) / sq           # greeks.py:21
    d2 = d1 - sq                                                       # greeks.py:22
    kd = K * mp.exp(-r * T)
    return S * ncdf(d1) - kd * ncdf(d2) if call else kd * ncdf(-d2) - S * ncdf(-d1)


def implied(price, S, K, T, r, call, guess):
    lp = mp.log(price)
    f = lambda s: mp.log(bs_price(S, K, T, r, s, call)) - lp      # on the logarithm: prices down to 1e-300 keep a relative tolerance
    lo, hi = mp.mpf(guess) / 2, mp.mpf(guess) * 2
    while f(lo) > 0:
        lo /= 2
    while f(hi) < 0:
        hi *= 2
    return mp.findroot(f, (lo, hi), solver="anderson", tol=mp.mpf(10) ** -40, maxsteps=200)


def iv_kats(n=2400, seed=2025):
    rng = np.random.default_rng(seed)
    rows = []
    while len(rows) < n:
        S = float(10.0 ** rng.uniform(-2, 5))
        m = float((10.0 ** rng.uniform(-7, -1) if rng.random() < 0.3 else rng.uniform(0.0, 3.0)) * rng.choice([-1.0, 1.0]))
        if len(rows) < 32:
            m = 0.0
        K = float(S * np.exp(m))
        T = float(10.0 ** rng.uniform(-4, 1.5))
        r = 0.0 if len(rows) < 16 else float(rng.uniform(-0.02, 0.15))
        sigma = float(10.0 ** rng.uniform(-2, 0.7))
        call = bool(rng.random() < 0.5)
        Sm, Km, Tm, rm, sm = (mp.mpf(v) for v in (S, K, T, r, sigma))
        p = float(bs_price(Sm, Km, Tm, rm, sm, call))
        kd = Km * mp.exp(-rm * Tm)
        intrinsic = max(Sm - kd, 0) if call else max(kd - Sm, 0)
        upper = Sm if call else kd
        tv = mp.mpf(p) - intrinsic
        if not (tv > mp.mpf(10) ** -300 and mp.mpf(p) < upper) or p <= 0.0:
            continue                                    # outside the invertible domain once the price is a double
        if tv < mp.mpf(p) * mp.mpf(2) ** -40 and intrinsic > 0:
            continue                                    # time value below the price's last bits: the double price says nothing about sigma
        try:
            root = implied(mp.mpf(p), Sm, Km, Tm, rm, call, sigma)
        except Exception:
            continue
        d1 = (mp.log(Sm / Km) + (rm + root * root / 2) * Tm) / (root * mp.sqrt(Tm))
        vega = Sm * mp.exp(-d1 * d1 / 2) / mp.sqrt(2 * mp.pi) * mp.sqrt(Tm)
        otm = (Sm < kd) if call else (Sm > kd)
        # d sigma per RELATIVE change of what the solver inverts: the price itself out of the money, the time value (a difference of
        # the price and max(S, K e^{-rT}) - min) in the money — 50-digit ratio, no underflow in the wings
        rows.append(dict(S=S, K=K, T=T, r=r, is_call=int(call), price=p, sigma=float(root), sigma_true=sigma, vega=float(vega), otm=int(otm),
                         dsigma_per_rel=float(min((mp.mpf(p) if otm else max(mp.mpf(p), Sm, Km)) / vega, mp.mpf(10) ** 300))))
    return rows


def spline_kat(x, y, xq):
    n = len(x)
    X, Y = [mp.mpf(v) for v in x], [mp.mpf(v) for v in y]
    M = [mp.mpf(0)] * n
    if n > 2:
        cp, dp = [mp.mpf(0)] * n, [mp.mpf(0)] * n
        for i in range(1, n - 1):
            hm, hp = X[i] - X[i - 1], X[i + 1] - X[i]
            den = 2 * (hm + hp) - hm * cp[i - 1]
            cp[i] = hp / den
            dp[i] = (6 * ((Y[i + 1] - Y[i]) / hp - (Y[i] - Y[i - 1]) / hm) - hm * dp[i - 1]) / den
        for i in range(n - 2, 0, -1):
            M[i] = dp[i] - cp[i] * M[i + 1]
    out = []
    for q in xq:
        v = mp.mpf(q)
        i = 0
        while i < n - 2 and v >= X[i + 1]:
            i += 1
        h, A, B = X[i + 1] - X[i], X[i + 1] - v, v - X[i]
        out.append(float((M[i] * A ** 3 + M[i + 1] * B ** 3) / (6 * h) + (Y[i] / h - M[i] * h / 6) * A + (Y[i + 1] / h - M[i + 1] * h / 6) * B))
    return [float(m) for m in M], out


def spline_kats(seed=77):
    rng = np.random.default_rng(seed)
    cases = []
    for n in (2, 3, 4, 5, 8, 17, 33, 64, 65, 100, 200, 333):
        for rep in range(2):
            x = np.cumsum(rng.uniform(0.05, 4.0, n) if rep else rng.uniform(0.9, 1.1, n)) + 2000.0
            y = 0.2 + 0.3 * np.log(x / (x.mean())) ** 2 + 0.004 * rng.standard_normal(n)
            xq = np.sort(np.concatenate([rng.uniform(x[0] - 5, x[-1] + 5, 40), x[:: max(1, n // 6)], [x[0], x[-1]]]))
            M, yq = spline_kat(x, y, xq)
            cases.append(dict(x=x.tolist(), y=y.tolist(), xq=xq.tolist(), M=M, yq=yq))
    return cases


if __name__ == "__main__":
    iv = iv_kats()
    with open(os.path.join(OUT, "iv_kats.json"), "w") as f:
        json.dump({"how": "tools/gen_kats_mpmath.py, mpmath %s at 50 digits" % mp.__version__, "quotes": iv}, f)
    sp = spline_kats()
    with open(os.path.join(OUT, "spline_kats.json"), "w") as f:
        json.dump({"how": "tools/gen_kats_mpmath.py, mpmath %s at 50 digits; extrapolation = polynomial continuation of the end segments" % mp.__version__,
                   "cases": sp}, f)
    print(f"{len(iv)} implied-vol quotes, {len(sp)} spline cases")
