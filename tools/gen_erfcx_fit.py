#!/usr/bin/env python
"""Coefficients of the erfcx fits in iv_interpolation_b200/csrc/ivx_bsiv.cu (mpmath, 50 digits).

    erfcx(q) = g(t) / (q + c),   t = (q - c)/(q + c) = 1 - 2c/(q + c)  in [-1, 1)  for q >= 0,
    g(t) = (q + c) erfcx(q)  ->  1/sqrt(pi) as q -> inf (analytic at t = 1).

ONE reciprocal 1/(q + c) gives both the argument t and the prefactor (the earlier fit of (1 + 2q) erfcx(q) needed 1/(q+4) AND
1/(1+2q): ten more fp64 operations per pair).  g is interpolated at Chebyshev nodes and converted to the monomial basis
(Horner in t, highest power first); the script prints the maximum relative error over q in [0, 1e3].

    python tools/gen_erfcx_fit.py            # fp64: c = 4, degree 19;  fp32: c = 2, degree 9
"""
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 50


def erfcx(q):
    q = mp.mpf(q)
    return mp.exp(q * q) * mp.erfc(q) if q < 20 else sum((-1) ** k * mp.rf(mp.mpf(1) / 2, k) / q ** (2 * k + 1) for k in range(40)) / mp.sqrt(mp.pi)


def g_of_t(t, c):
    t = mp.mpf(t)
    if t >= 1:
        return 1 / mp.sqrt(mp.pi)
    q = c * (1 + t) / (1 - t)
    return (q + c) * erfcx(q)


def monomial_fit(c, deg):
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
    vals = [g_of_t(t, c) for t in nodes]
    cheb = [sum(vals[i] * mp.cos(mp.pi * j * (2 * i + 1) / (2 * n)) for i in range(n)) * 2 / n for j in range(n)]
    cheb[0] /= 2
    # Chebyshev -> monomial: T_0 = 1, T_1 = t, T_{k+1} = 2 t T_k - T_{k-1}
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for k in range(1, deg):
        nxt = [mp.mpf(0)] + [2 * a for a in T[k]]
        for i, a in enumerate(T[k - 1]):
            nxt[i] -= a
        T.append(nxt)
    mono = [mp.mpf(0)] * n
    for j in range(n):
        for i, a in enumerate(T[j]):
            mono[i] += cheb[j] * a
    return mono[::-1]                     # highest power first


def max_rel_err(coef, c, dtype):
    qs = np.concatenate([np.linspace(0, 30, 1500), np.logspace(-8, 3, 700)])
    worst = 0.0
    for q in qs:
        qd = dtype(q)
        r = dtype(1) / (qd + dtype(c))
        t = dtype(1) - dtype(2 * c) * r
        p = dtype(float(coef[0]))
        for a in coef[1:]:
            p = p * t + dtype(float(a))
        got = float(p * r)
        ex = float(erfcx(float(qd)))
        worst = max(worst, abs(got / ex - 1))
    return worst


if __name__ == "__main__":
    for name, c, deg, dtype, fmt in (("ERFCX_C (fp64)", 4, 19, np.float64, "{:.17e}"), ("erfcx_pos_f (fp32)", 2, 9, np.float32, "{:.9e}f")):
        coef = monomial_fit(c, deg)
        print(f"// {name}: c = {c}, degree {deg}, max relative error {max_rel_err(coef, c, dtype):.1e} (evaluated in {dtype.__name__})")
        print(",\n".join("    " + fmt.format(float(a)) for a in coef))
