"""How well a double-precision option price determines its implied volatility — the explicit per-quote criterion the IV tests use
instead of quotas:  |sigma - sigma_ref| <= 1e-8 + ulps x 2^-52 x (inverted quantity / vega), where the inverted quantity is the
price itself out of the money and max(price, S, K) in the money (the time value is then a difference of numbers of that size).
Works on numpy arrays and on torch tensors (the full-size tests keep everything on the device)."""
import math


def iv_allowed(price, S, K, T, r, is_call, sigma, ulps=64.0, xp=None):
    if xp is None:
        import numpy as xp
    call = is_call != 0
    sq = sigma * xp.sqrt(T)
    d1 = (xp.log(S / K) + (r + 0.5 * sigma * sigma) * T) / sq
    ln_vega = xp.log(S) - 0.5 * d1 * d1 - 0.5 * math.log(2.0 * math.pi) + 0.5 * xp.log(T)
    kd = K * xp.exp(-r * T)
    otm = xp.where(call, S < kd, S > kd)
    big = xp.maximum(xp.maximum(price, S), K)
    ln_q = xp.log(xp.where(otm, price, big))
    return 1e-8 + ulps * 2.0 ** -52 * xp.exp(xp.clip(ln_q - ln_vega, -700.0, 700.0))
