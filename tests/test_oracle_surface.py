"""CPU tests of the surface-pipeline oracle (north_star kernels C/D: no reference code, scipy is the anchor)."""
import numpy as np

from iv_interpolation_b200 import tapes
from oracle import np_oracle as O, surface_oracle as SO


def test_fast_implied_vol_matches_brentq():
    q = tapes.make_quote_tape(6, 3000, pricer=O.bs_price)
    price = q.price.copy()
    price[:6] = [np.nan, -1.0, 0.0, np.inf, 1e9, q.S[5] * 2]
    a = O.implied_vol(price, q.S, q.K, q.T, q.r, q.is_call)
    b = SO.implied_vol_fast(price, q.S, q.K, q.T, q.r, q.is_call)
    assert (np.isnan(a) == np.isnan(b)).all()
    ok = ~np.isnan(a)
    with np.errstate(all="ignore"):
        kd = q.K * np.exp(-q.r * q.T)
        tv = price - np.maximum(np.where(q.is_call != 0, q.S - kd, kd - q.S), 0.0)
    well = ok & (tv > 1e-6 * q.S)          # the price still determines sigma to 1e-8
    assert well.sum() > 2000
    assert np.abs(a - b)[well].max() <= 1e-9
    # everywhere else both must at least reprice the quote to rounding
    rep = O.bs_price(q.S[ok], q.K[ok], q.T[ok], q.r[ok], b[ok], q.is_call[ok])
    assert (np.abs(rep - price[ok]) <= 64 * np.spacing(np.maximum(price[ok], q.S[ok]))).all()


def test_pipeline_round_trip_and_shapes():
    c = tapes.make_quote_chain_np(3, 130, 4, 40, O.bs_price, n_grid_k=24, n_grid_m=8, drop_frac=0.02)
    res = SO.surface_pipeline(c["ts"], c["price"], c["S"], c["K"], c["T"], c["r"], c["is_call"], 4, 40, c["grid_k"], c["grid_m"],
                              60 * tapes.US, 30)
    iv = res["iv"]
    both = ~np.isnan(iv)
    assert abs(np.isnan(iv).mean() - 0.02) < 0.01
    with np.errstate(all="ignore"):
        kd = c["K"] * np.exp(-c["r"] * c["T"])
        tv = c["price"] - np.maximum(np.where(c["is_call"] != 0, c["S"] - kd, kd - c["S"]), 0.0)
    well = both & (tv > 1e-6 * c["S"])
    assert np.abs(iv - c["sigma_true"])[well].max() <= 1e-8          # sigma -> price -> sigma
    assert res["smile"].shape == (130, 4, 24) and res["surface"].shape == (130, 8, 24)
    # 09:30:00 + 130 s -> buckets of 60, 60, 10 stamps; the last one is incomplete at min_count=30
    assert res["count"].tolist() == [60, 60, 10]
    assert np.isnan(res["open"][2]).all() and not np.isnan(res["open"][:2]).any()
    flat = res["surface"].reshape(130, -1)
    assert (res["open"][0] == flat[0]).all() and (res["close"][1] == flat[119]).all()
    assert (res["high"][0] == flat[:60].max(axis=0)).all() and (res["low"][1] == flat[60:120].min(axis=0)).all()
    # candle rules agree with the reference-pinned kernel-B oracle on one series
    s = 5
    w = O.ohlcv_reduce(c["ts"], flat[:, s], flat[:, s], flat[:, s], flat[:, s], np.zeros(130), bucket_ticks=60 * tapes.US, min_count=1)
    for k in ("open", "high", "low", "close"):
        assert (w[k][:2] == res[k][:2, s]).all()
