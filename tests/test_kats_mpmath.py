"""Known-answer vectors computed with mpmath at 50 digits (tools/gen_kats_mpmath.py -> tests/golden/iv_kats.json,
spline_kats.json): the oracles of the two kernels the reference has no code for stop being "whatever scipy is installed".
CPU tier: the oracles against the vectors.  GPU tier: the kernels against the vectors, through the C ABI.

Implied vol: |sigma - sigma_50digit| <= 1e-8 + 32 ulp x (inverted quantity / vega) — the forward bound of the north_star, plus what the
double-precision price itself cannot resolve (an explicit conditioning criterion per quote, no quota)."""
import json
import os

import numpy as np
import pytest

from oracle import np_oracle as O, surface_oracle as SO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def iv_kats():
    with open(os.path.join(GOLDEN, "iv_kats.json")) as f:
        q = json.load(f)["quotes"]
    return {k: np.array([r[k] for r in q], dtype=np.uint8 if k == "is_call" else np.float64) for k in q[0]}


@pytest.fixture(scope="module")
def spline_kats():
    with open(os.path.join(GOLDEN, "spline_kats.json")) as f:
        return json.load(f)["cases"]


def iv_allowed(k):
    """1e-8 + what 32 ulps of the inverted quantity (the price out of the money; max(price, S, K) in the money, where the time
    value is a difference) move sigma by: d sigma = relative change x (that quantity / vega), the ratio taken at 50 digits."""
    return 1e-8 + 32 * 2.0 ** -52 * k["dsigma_per_rel"]


def test_iv_oracles_match_the_50_digit_vectors(iv_kats):
    k = iv_kats
    assert len(k["S"]) >= 2000 and (k["dsigma_per_rel"] > 0).all()
    assert (iv_allowed(k) <= 2e-8).mean() > 0.6, "most vectors must be well conditioned: the forward bound itself applies to them"
    fast = SO.implied_vol_fast(k["price"], k["S"], k["K"], k["T"], k["r"], k["is_call"])
    assert not np.isnan(fast).any()
    err = np.abs(fast - k["sigma"])
    worst = int(np.argmax(err - iv_allowed(k)))
    assert (err <= iv_allowed(k)).all(), (err[worst], iv_allowed(k)[worst], {n: k[n][worst] for n in ("S", "K", "T", "r", "sigma", "price")})
    sub = slice(0, 400)                                    # brentq per quote: a sample
    br = O.implied_vol(k["price"][sub], k["S"][sub], k["K"][sub], k["T"][sub], k["r"][sub], k["is_call"][sub])
    assert (np.abs(br - k["sigma"][sub]) <= iv_allowed(k)[sub]).all()
    # the vectors cover the regimes: deep wings, short and long maturities, low and high vols, both sides of the money
    m = np.log(k["K"] / k["S"])
    assert (np.abs(m) > 2).sum() > 50 and (np.abs(m) < 1e-3).sum() > 100 and (k["T"] < 1e-3).sum() > 100 and (k["sigma"] > 2).sum() > 100


def test_spline_oracle_matches_the_50_digit_vectors(spline_kats):
    for c in spline_kats:
        x, y, xq = (np.array(c[n]) for n in ("x", "y", "xq"))
        got = O.natural_spline_eval(x, y, xq, 1)
        scale = max(1.0, np.abs(c["yq"]).max())
        assert np.abs(got - np.array(c["yq"])).max() <= 1e-10 * scale, (len(x), np.abs(got - np.array(c["yq"])).max())
        M = O.natural_spline_moments(x, y)
        assert np.abs(M - np.array(c["M"])).max() <= 1e-9 * max(1.0, np.abs(c["M"]).max()), len(x)


@pytest.mark.gpu
@pytest.mark.parametrize("two_kernels", [True, False])
def test_iv_kernel_matches_the_50_digit_vectors(iv_kats, two_kernels):
    import torch
    from iv_interpolation_b200 import device as D
    dev = torch.device("cuda:0")
    k = iv_kats
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sig = D.bs_implied_vol(t(k["price"]), t(k["S"]), t(k["K"]), t(k["T"]), t(k["r"]), t(k["is_call"]), two_kernels=two_kernels).cpu().numpy()
    assert not np.isnan(sig).any()
    err = np.abs(sig - k["sigma"])
    worst = int(np.argmax(err - iv_allowed(k)))
    assert (err <= iv_allowed(k)).all(), (err[worst], iv_allowed(k)[worst], {n: k[n][worst] for n in ("S", "K", "T", "r", "sigma", "price")})
    # chain layout: the same quotes as one stamp x one expiry x n strikes is not possible (S, T, r differ per quote); one quote per chain
    n = 64
    for i in range(0, n):
        one = D.bs_implied_vol_chain(t(k["price"][i:i + 1]), t(k["is_call"][i:i + 1]), t(k["S"][i:i + 1]), t(k["K"][i:i + 1]), t(k["T"][i:i + 1]),
                                     t(k["r"][i:i + 1]), n_t=1, n_exp=1, n_strike=1).cpu().numpy()
        assert abs(one[0] - k["sigma"][i]) <= iv_allowed(k)[i], i


@pytest.mark.gpu
def test_spline_kernels_match_the_50_digit_vectors(spline_kats):
    import torch
    from iv_interpolation_b200 import device as D
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for c in spline_kats:
        x, y, xq = (np.array(c[n]) for n in ("x", "y", "xq"))
        off = np.array([0, len(x)], np.int64)
        got = D.spline_fit_eval(t(off), t(x), t(y), t(xq), max_knots=len(x), extrap=1).cpu().numpy()
        scale = max(1.0, np.abs(c["yq"]).max())
        assert np.abs(got - np.array(c["yq"])).max() <= 1e-9 * scale, (len(x), np.abs(got - np.array(c["yq"])).max())
        if len(x) <= 64:                                   # the maturity-direction kernel: [knot][system] planes
            got2 = D.spline_strided(t(x), t(np.repeat(y, 3)), t(xq), 3, extrap=1).cpu().numpy().reshape(len(xq), 3)
            assert np.abs(got2 - np.array(c["yq"])[:, None]).max() <= 1e-9 * scale, len(x)
