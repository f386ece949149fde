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
