"""The batched host path (iv_interpolation_b200/hostpath.py: concat -> one launch -> one result frame -> slices) on the CPU tier:
the launch is replaced by the oracle-backed stand-in of tests/fake_device.py, everything else is the product's code.  Held to the
reference-generated goldens and, where the reference is present (/root/reference or oracle/_ref), to the reference's classes on
random batches — frames in, frames out, bit for bit, index labels and dtypes included."""
import json

import numpy as np
import pandas as pd
import pytest
import torch

from iv_interpolation_b200 import hostpath as H, tapes
from oracle import golden_io, reference_loader as RL
from tests import fake_device as FD

CPU = torch.device("cpu")


def _uniform_interp_cases(golden_interp):
    names = json.loads(str(golden_interp["names"]))
    out = []
    for n in names:
        f = golden_io.dict_to_frame(f"{n}/in", golden_interp)
        kw = json.loads(str(golden_interp[f"{n}/kwargs"]))
        if kw.get("method", "linear") == "linear" and f is not None and H.task1_uniform([f]) and len(f) >= kw.get("min_points", 10):
            out.append((n, f, kw.get("min_points", 10)))
    return out


def test_interpolate_uniform_matches_reference_goldens(golden_interp):
    cases = _uniform_interp_cases(golden_interp)
    assert len(cases) >= 20
    by_mp = {}
    for n, f, mp in cases:
        by_mp.setdefault(mp, []).append((n, f))
    for mp, group in by_mp.items():
        sigs = {}
        for n, f in group:                                       # one batch per common schema
            sigs.setdefault((tuple(f.columns), tuple(map(str, f.dtypes))), []).append((n, f))
        for members in sigs.values():
            outs = H.interpolate_uniform([f for _, f in members], mp, launch=FD.interp_linear_ffill, device=CPU)
            for (n, _), got in zip(members, outs):
                try:
                    golden_io.assert_frames_match(golden_io.dict_to_frame(f"{n}/out", golden_interp), got)
                except AssertionError as e:
                    raise AssertionError(f"case {n}: {e}") from None


def test_reconstruct_uniform_matches_reference_goldens(golden_candles):
    names = json.loads(str(golden_candles["names"]))
    by_freq = {}
    for n in names:
        f = golden_io.dict_to_frame(f"{n}/in", golden_candles)
        if f is not None and not f.empty and H.task2_uniform([f]) and not f["symbol"].isna().any():
            by_freq.setdefault(str(golden_candles[f"{n}/freq"]), []).append((n, f))
    n_checked = 0
    for freq, group in by_freq.items():
        nmin = int(freq[:-3]) if freq.endswith("min") else int(freq[:-1])
        if nmin < 1:
            continue
        sigs = {}
        for n, f in group:
            sigs.setdefault((tuple(f.columns), tuple(map(str, f.dtypes))), []).append((n, f))
        for members in sigs.values():
            got = H.reconstruct_uniform([f for _, f in members], nmin, freq, pd.Timestamp("2024-01-01"), presorted=False, check_len=True,
                                        launch=FD.ohlcv_reduce, device=CPU)
            assert got is not None
            for (n, _), g in zip(members, got[0]):
                try:
                    golden_io.assert_frames_match(golden_io.dict_to_frame(f"{n}/out", golden_candles), g)
                except AssertionError as e:
                    raise AssertionError(f"case {n} ({freq}): {e}") from None
                n_checked += 1
    assert n_checked >= 20


@pytest.mark.skipif(not RL.available(), reason="neither /root/reference nor oracle/_ref is present")
def test_batched_host_path_equals_the_reference_class_loop():
    """A config-1-shaped batch (contracts x 25 hourly rows), some frames shuffled in time, through the batched path and through
    the reference's own IVInterpolator / CandleReconstructor one frame at a time."""
    tp = tapes.make_contract_tape(17, 40)
    rng = np.random.default_rng(3)
    frames = []
    for i in range(tp.n_sym):
        f = tp.frame(i)
        if i % 5 == 2:
            f = f.iloc[rng.permutation(len(f))].reset_index(drop=True)
        if i % 9 == 4:
            f.loc[0, "callput"] = None
        frames.append(f)
    assert H.task1_uniform(frames)
    outs = H.interpolate_uniform(frames, 10, launch=FD.interp_linear_ffill, device=CPU)
    for f, got in zip(frames, outs):
        golden_io.assert_frames_match(RL.reference_interpolate(f), got)
    minute = []
    for got in outs[:12]:
        m = pd.DataFrame({"symbol": got["symbol"].to_numpy(), "timestamp": got["date"].to_numpy(), "open": got["underlying_price"].to_numpy(),
                          "high": got["underlying_price"].to_numpy() + 3.0, "low": got["underlying_price"].to_numpy() - 3.0,
                          "close": got["underlying_price"].to_numpy() + 0.5, "volume": got["volume"].to_numpy()})
        minute.append(m)
    minute[3] = minute[3].iloc[::-1].reset_index(drop=True)
    assert H.task2_uniform(minute)
    res, flags = H.reconstruct_uniform(minute, 5, "5min", pd.Timestamp("2024-01-01"), presorted=False, check_len=True,
                                       launch=FD.ohlcv_reduce, device=CPU)
    for m, got in zip(minute, res):
        golden_io.assert_frames_match(RL.reference_reconstruct(m, "5min"), got)


def test_uniform_schema_detection():
    tp = tapes.make_contract_tape(1, 3)
    a, b = tp.frame(0), tp.frame(1)
    assert H.task1_uniform([a, b])
    assert not H.task1_uniform([a, b.assign(extra=1.0)])                      # extra column -> general path
    assert not H.task1_uniform([a, b.astype({"strike": np.int64})])           # integer numeric -> general path
    assert not H.task1_uniform([a, b[list(b.columns[::-1])]])                  # another column order
    assert not H.task1_uniform([a.assign(date=a["date"].dt.tz_localize("UTC"))])
    assert not H.task1_uniform([])


@pytest.mark.skipif(not RL.available(), reason="neither /root/reference nor oracle/_ref is present")
def test_integer_fill_columns_take_the_numeric_path_like_the_reference():
    """Integer strike / volume columns (and five non-float fill columns in all) used to overflow the four string slots of a
    launch; they travel as float64 (exact below 2^53) and come back float64 — or with their own dtype where the grid is dense and
    the reference's merge introduces no NaN.  Host logic shared by the product and the DataFrame oracle, held to the reference."""
    from iv_interpolation_b200 import frames as F
    from oracle.df_oracle import OracleIVInterpolator
    tp = tapes.make_contract_tape(1, 2, rows_per_contract=14)
    f = tp.frame(0)
    f["strike"] = f["strike"].astype(np.int64)
    f["volume"] = np.arange(len(f), dtype=np.int64)
    inp = F.interp_frame_to_arrays(f)
    assert inp.fill_idx_names == ["symbol", "callput", "record_time"] and set(inp.fill_int_dtype) == {"strike", "volume"}
    golden_io.assert_frames_match(RL.reference_interpolate(f), OracleIVInterpolator().interpolate_symbol(f))
    g = tp.frame(1)                                              # the same on a dense 1-minute grid: dtypes survive
    g["date"] = pd.date_range("2023-03-15", periods=len(g), freq="1min")
    g["strike"] = g["strike"].astype(np.int64)
    g["quote_volume"] = np.arange(len(g), dtype=np.int32)
    want, got = RL.reference_interpolate(g), OracleIVInterpolator().interpolate_symbol(g)
    assert want["strike"].dtype == np.int64 and got["strike"].dtype == np.int64 and got["quote_volume"].dtype == np.int32
    golden_io.assert_frames_match(want, got)


@pytest.mark.skipif(not RL.available(), reason="neither /root/reference nor oracle/_ref is present")
def test_tz_aware_candles_floor_the_local_wall_clock():
    """dt.floor on a tz-aware column floors local wall time: 60-minute candles in +05:30 start on the local hour, not the UTC one;
    a leading partial bucket (index labels 1..) must not misalign the stamps.  Frames that straddle a DST change are refused."""
    from oracle.df_oracle import OracleCandleReconstructor
    n = 200
    b = 100 + np.cumsum(np.random.default_rng(0).normal(0, 1, n))
    for tz in ("Asia/Kolkata", "UTC", "America/New_York"):
        ts = pd.date_range("2024-03-01 09:07", periods=n, freq="1min", tz=tz)
        f = pd.DataFrame({"symbol": "S", "timestamp": ts, "open": b, "high": b + 1, "low": b - 1, "close": b + 0.2, "volume": 1.0})
        for freq in ("60min", "5min", "45min"):
            golden_io.assert_frames_match(RL.reference_reconstruct(f, freq), OracleCandleReconstructor(freq).reconstruct_symbol_candles(f))
    dst = pd.date_range("2024-03-10 01:30", periods=90, freq="1min", tz="America/New_York")      # crosses 02:00 -> 03:00
    f = pd.DataFrame({"symbol": "S", "timestamp": dst, "open": b[:90], "high": b[:90] + 1, "low": b[:90] - 1, "close": b[:90], "volume": 1.0})
    assert OracleCandleReconstructor("5min").reconstruct_symbol_candles(f) is None
    # Task 1 with tz-aware dates and dropped leading rows (index labels start above 0)
    from oracle.df_oracle import OracleIVInterpolator
    tp = tapes.make_contract_tape(3, 8)
    g = tp.frame(3)
    g["date"] = g["date"].dt.tz_localize("Asia/Kolkata")
    g.loc[0, "iv"] = np.nan
    golden_io.assert_frames_match(RL.reference_interpolate(g), OracleIVInterpolator().interpolate_symbol(g))


def test_sparse_underlying_tape_is_a_consistent_ragged_csr():
    """tapes.make_underlying_tape(drop_frac=...): the replay tape of BASELINE configs[4] — hourly rows missing at random, never a
    contract's first or last; offsets, stamps and columns stay aligned and every contract stays time-sorted."""
    import numpy as np
    from iv_interpolation_b200 import tapes
    full = tapes.make_underlying_tape(7, 40, 121, nan_frac=0.15)
    tp = tapes.make_underlying_tape(7, 40, 121, nan_frac=0.15, drop_frac=0.25)
    n = np.diff(tp.off)
    assert tp.off[0] == 0 and tp.off[-1] == len(tp.date) and (n >= 2).all() and (n <= 121).all() and 0.6 < n.mean() / 121 < 0.9
    assert all(len(c) == len(tp.date) for c in tp.cols.values()) and len(tp.symbols) == 40
    for i in range(40):
        a, b = int(tp.off[i]), int(tp.off[i + 1])
        fa, fb = int(full.off[i]), int(full.off[i + 1])
        assert (np.diff(tp.date[a:b]) > 0).all()
        assert tp.date[a] == full.date[fa] and tp.date[b - 1] == full.date[fb - 1]          # first and last row kept
        assert np.isin(tp.date[a:b], full.date[fa:fb]).all()
    assert np.isnan(tp.cols["iv"]).mean() > 0.1


def test_pipeline_batch_whole_frame_handover_equals_per_symbol_handover(monkeypatch):
    """PipelineBatch hands each stage's whole result frame to the next stage (hostpath.FrameSlices) instead of per-symbol frames
    that are cut, column-selected and concatenated again.  Both hand-overs must give the same frames at every stage (values,
    dtypes, index labels) and the same status dictionaries — with skipped symbols, a symbol below min_points and one whose
    interpolation fails in the batch."""
    from iv_interpolation_b200 import _lib, device as D
    from iv_interpolation_b200.batch import PipelineBatch

    monkeypatch.setattr(D, "interp_linear_ffill", FD.interp_linear_ffill)
    monkeypatch.setattr(D, "ohlcv_reduce", FD.ohlcv_reduce)
    monkeypatch.setattr(D, "bridge_plan", FD.bridge_plan)
    monkeypatch.setattr(D, "bridge_ohlcv", FD.bridge_ohlcv)
    monkeypatch.setattr(H, "_to_device", lambda a, dev: torch.from_numpy(np.array(a, copy=True)))
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(_lib, "lib", lambda: None)                           # the entries check that libivx.so loads before touching data

    def batch(shape):
        tp = tapes.make_contract_tape(5, 10, rows_per_contract=12)         # 12 hourly rows -> 661 minute rows each
        frames = {f"c{i}": tp.frame(i) for i in range(tp.n_sym)}
        if shape == "ragged":
            frames["c2"] = frames["c2"].iloc[:0]                             # empty -> skipped at every stage
            frames["c4"] = frames["c4"].iloc[:2]                             # below min_points
            f = frames["c6"].copy()
            f["iv"] = np.nan                                                 # nothing left after dropna -> None
            frames["c6"] = f
        return frames

    for shape in ("clean", "ragged"):
        frames = batch(shape)
        run = lambda whole: PipelineBatch(rng=np.random.RandomState(2), batch_id=7, whole_frames=whole).run(
            {k: v for k, v in frames.items()})
        built = []
        init = H.FrameSlices.__init__
        monkeypatch.setattr(H.FrameSlices, "__init__", lambda self, big, off: (built.append(len(off) - 1), init(self, big, off))[1])
        a = run(True)
        assert built == [len(a[0]), len(a[1])]                               # both hand-overs went by the whole frame ...
        b = run(False)
        assert len(built) == 2                                               # ... and none of them here
        monkeypatch.setattr(H.FrameSlices, "__init__", init)
        assert len(a[0]) == (10 if shape == "clean" else 7)
        for k in range(3):
            assert list(a[k]) == list(b[k])
            for s in a[k]:
                pd.testing.assert_frame_equal(a[k][s], b[k][s], check_exact=True)
        strip = lambda L: [{k: v for k, v in d.items() if k != "processing_time"} for d in L]
        for name in a[3]:
            assert strip(a[3][name][0]) == strip(b[3][name][0])
