"""The numpy oracle is held, bit for bit, to golden vectors produced by the reference's own
classes (oracle/make_golden.py).  CPU only."""
import json

import numpy as np
import pytest

from oracle import golden_io, np_oracle as O
from oracle.df_oracle import OracleCandleReconstructor, OracleIVInterpolator


def _names(npz):
    return json.loads(str(npz["names"]))


def test_interp_oracle_matches_reference_golden(golden_interp):
    names = _names(golden_interp)
    assert len(names) >= 50
    n_frames = 0
    for name in names:
        frame = golden_io.dict_to_frame(f"{name}/in", golden_interp)
        want = golden_io.dict_to_frame(f"{name}/out", golden_interp)
        kw = json.loads(str(golden_interp[f"{name}/kwargs"]))
        got = OracleIVInterpolator(**kw).interpolate_symbol(frame)
        try:
            golden_io.assert_frames_match(want, got)
        except AssertionError as e:
            raise AssertionError(f"case {name}: {e}") from None
        n_frames += want is not None
    assert n_frames >= 40


def test_candle_oracle_matches_reference_golden(golden_candles):
    names = _names(golden_candles)
    assert len(names) >= 40
    for name in names:
        frame = golden_io.dict_to_frame(f"{name}/in", golden_candles)
        want = golden_io.dict_to_frame(f"{name}/out", golden_candles)
        freq = str(golden_candles[f"{name}/freq"])
        rec = OracleCandleReconstructor(freq)
        got = rec.reconstruct_symbol_candles(frame)
        try:
            golden_io.assert_frames_match(want, got)
            if all(c in frame.columns for c in ("open", "high", "low", "close", "volume")):
                assert rec.validate_candle_data(frame) == bool(golden_candles[f"{name}/valid_in"])
            if want is not None:
                assert rec.validate_candle_data(want) == bool(golden_candles[f"{name}/valid_out"])
        except AssertionError as e:
            raise AssertionError(f"case {name}: {e}") from None


def test_survey_known_answers(golden_interp, golden_candles):
    """Spot values quoted in SURVEY.md §8(c), read back from the reference-generated goldens."""
    a1 = golden_io.dict_to_frame("A1_hourly3/out", golden_interp)
    assert len(a1) == 121 and a1["iv"].iloc[30] == 0.65 and a1["underlying_price"].iloc[30] == 100.5
    assert a1["time_to_maturity"].iloc[30] == 0.975 and a1["record_time"].iloc[30] == "r0"
    assert not a1["is_interpolated"].any()
    a5 = golden_io.dict_to_frame("A5_duplicate_stamp/out", golden_interp)
    assert np.allclose(a5["iv"], [.5, .6, .7, .9, 1.0, 1.1]) and len(a5) == 6
    a6 = golden_io.dict_to_frame("A6_nan_lead_mid_tail/out", golden_interp)
    assert np.allclose(a6["iv"], [.4, .5, .6, .7, .8, .8, .8]) and list(a6.index) == list(range(2, 9))
    assert golden_io.dict_to_frame("A7_span_31d_none/out", golden_interp) is None
    assert len(golden_io.dict_to_frame("span_exactly_30d/out", golden_interp)) == 43201
    b1 = golden_io.dict_to_frame("B1_23rows_from_0902/out", golden_candles)
    assert list(b1.index) == [1, 2, 3, 4]
    assert b1[["open", "high", "low", "close", "volume"]].iloc[0].tolist() == [103.0, 108.0, 102.0, 107.5, 5.0]
    b6 = golden_io.dict_to_frame("B6_all_nan_open_volume/out", golden_candles)
    assert np.isnan(b6["open"].iloc[0]) and b6["volume"].iloc[0] == 0.0


def test_parse_frequency_kat(golden_kats):
    for fr, want in golden_kats["parse_frequency"].items():
        if want == "ValueError":
            with pytest.raises(ValueError):
                OracleCandleReconstructor(fr)
        else:
            assert OracleCandleReconstructor(fr).frequency_minutes == want


def test_greeks_kat(golden_kats):
    for case in golden_kats["greeks"]:
        c = dict(case["in"])
        call = c.pop("option_type") == "call"
        got = O.bs_greeks(c["S"], c["K"], c["T"], c["r"], c["sigma"], is_call=call)
        for k, v in case["out"].items():
            assert abs(float(got[k]) - v) <= 1e-13 * max(1.0, abs(v)), (k, float(got[k]), v)


def test_interp_restated_is_np_interp():
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 40))
        xp = np.sort(rng.choice(2000, n, replace=False)).astype(float)
        fp = rng.normal(size=n) * 10.0 ** rng.integers(-3, 4)
        if n > 2 and rng.random() < 0.3:
            fp[rng.integers(0, n)] = rng.choice([np.inf, -np.inf])
        x = rng.uniform(-10, 2010, 300).round()
        with np.errstate(all="ignore"):
            want = np.interp(x, xp, fp)
        got = O.interp_restated(x, xp, fp)
        assert ((want == got) | (np.isnan(want) & np.isnan(got))).all()


def test_kahan_differs_from_naive_and_matches_golden(golden_candles):
    f = golden_io.dict_to_frame("B5_kahan_2000_buckets/in", golden_candles)
    want = golden_io.dict_to_frame("B5_kahan_2000_buckets/out", golden_candles)
    v = f["volume"].to_numpy()
    naive = np.array([np.add.reduce(v[i:i + 5]) if False else sum(v[i:i + 5].tolist()) for i in range(0, len(v), 5)])
    kahan = O.kahan_group_sum(v, np.arange(0, len(v), 5), np.arange(5, len(v) + 5, 5))
    assert (kahan == want["volume"].to_numpy()).all()
    assert (naive != kahan).sum() > 100          # compensation matters: a naive sum is NOT bit-equal


def test_bridge_oracle_matches_reference_golden(golden_bridge):
    """oracle/bridge_oracle.py against candles produced by the reference's own data bridge under a seeded generator
    (oracle/make_golden_bridge.py): every strategy, NaN / zero / negative base prices and volumes, bit for bit."""
    from oracle import bridge_oracle as B
    man = json.loads(str(golden_bridge["manifest"]))
    assert len(man["cases"]) >= 36
    seen = set()
    for c in man["cases"]:
        k = c["key"]
        r = B.bridge_ohlcv(c["strategy"], golden_bridge[k + "/base"], golden_bridge[k + "/volume"],
                           golden_bridge[k + "/u"], golden_bridge[k + "/e"], golden_bridge[k + "/z"])
        assert (r["n_u"], r["n_e"], r["n_z"]) == tuple(len(golden_bridge[k + "/" + t]) for t in "uez"), c
        if k + "/none" in golden_bridge.files:
            assert len(r["row"]) == 0, c
            continue
        for col in ("open", "high", "low", "close", "volume", "source_price"):
            assert np.array_equal(golden_bridge[k + "/out_" + col], r[col]), (c, col)
        assert np.array_equal(golden_bridge[k + "/ts"][r["row"]], golden_bridge[k + "/out_ts"]), c
        if k + "/quality" in golden_bridge.files:
            assert B.quality_verdict(r["quality"]) == json.loads(str(golden_bridge[k + "/quality"])), c
        seen.add(B.canonical_strategy(c["strategy"]))
    assert seen == set(B.STRATEGIES)


def test_bridge_tapes_follow_the_generator():
    """draw_tapes consumes a RandomState exactly like row-by-row uniform/exponential/normal calls would."""
    from oracle import bridge_oracle as B
    rng = np.random.default_rng(5)
    base = 100 + rng.normal(0, 1, 200)
    base[::17] = np.nan
    vol = rng.exponential(3, 200)
    vol[::5] = np.nan
    for strat, k in (("spread_simulation", 5), ("pipeline", 4), ("trend_following", 0)):
        a, b = np.random.RandomState(9), np.random.RandomState(9)
        u, e, z = B.draw_tapes(a, strat, base, vol)
        uu, ee, zz = [], [], []
        for bp, v in zip(base, vol):
            if np.isnan(bp) or bp <= 0:
                continue
            if strat == "trend_following":
                zz.append(b.normal(0, 1))
            uu += [b.uniform(0, 1) for _ in range(k)]
            if np.isnan(v) or v <= 0:
                ee.append(b.exponential(1))
        assert np.array_equal(u, uu) and np.array_equal(e, ee) and np.array_equal(z, zz), strat
        assert a.random_sample() == b.random_sample()
