"""GPU parity of the three-stage contracts pipeline (iv_interpolation_b200/contracts.py: Task 1 -> data bridge -> Task 2, the order
reference complete_pipeline.py:740-831 runs them in) against the three reference-pinned oracle restatements chained contract after
contract, bit for bit; the device random tapes; the host-streaming entry against the resident one."""
import numpy as np
import pytest

from oracle import bridge_oracle as B, np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _t(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


def _oracle_chain(tp, i, rs, strategy):
    """Contract i through the three oracles; draws its bridge tapes from ``rs`` in the reference's order."""
    from iv_interpolation_b200 import tapes
    a, b = int(tp.off[i]), int(tp.off[i + 1])
    w1 = O.interp_linear_ffill(tp.date[a:b], [tp.cols[c][a:b] for c in tapes.INTERP_COLS], [tp.cols[c][a:b] for c in tapes.FILL_NUM_COLS],
                               [np.ones(b - a, bool)], step=tapes.MINUTE_US, max_span=30 * tapes.DAY_US)
    if w1["status"] != 0:
        return None, None, None, (np.zeros(0), np.zeros(0))
    base, vol = w1["interp"][1], w1["fill"][tapes.FILL_NUM_COLS.index("volume")]
    u, e, z = B.draw_tapes(rs, strategy, base, vol)
    w2 = B.bridge_ohlcv(strategy, base, vol, u, e, z)
    ts = w1["date"][w2["row"]]
    w3 = O.ohlcv_reduce(ts, w2["open"], w2["high"], w2["low"], w2["close"], w2["volume"], bucket_ticks=5 * tapes.MINUTE_US, min_count=5)
    return w1, w2, w3, (u, e)


@pytest.mark.parametrize("strategy", ["spread_simulation", "pipeline", "simple_spread"])
def test_contract_pipeline_equals_the_oracle_chain(dev, strategy):
    from iv_interpolation_b200 import contracts as CT, tapes
    tp = tapes.make_contract_tape(31, 14, rows_per_contract=73)
    rs = np.random.RandomState(5)
    want = [_oracle_chain(tp, i, rs, strategy) for i in range(tp.n_sym)]
    u = np.concatenate([w[3][0] for w in want]); e = np.concatenate([w[3][1] for w in want] + [np.zeros(1)])
    pipe = CT.ContractPipeline(dev, strategy=strategy)
    res = pipe.run(_t(tp.off, dev), _t(tp.date, dev), [_t(tp.cols[c], dev) for c in tapes.INTERP_COLS],
                   [_t(tp.cols[c], dev) for c in tapes.FILL_NUM_COLS], tapes=(_t(u, dev), _t(e, dev)))
    h = lambda x: x.cpu().numpy()
    a_off, m_off, c_off = h(res.interp.out_off), h(res.minute.out_off), h(res.candles.out_off)
    n_ok = 0
    for i, (w1, w2, w3, _) in enumerate(want):
        if w1 is None:
            assert a_off[i + 1] == a_off[i] and c_off[i + 1] == c_off[i]
            continue
        sl = slice(a_off[i], a_off[i + 1])
        assert _eq(h(res.interp.ts)[sl], w1["date"]) and all(_eq(h(res.interp.interp[k])[sl], w1["interp"][k]) for k in range(3))
        sm = slice(m_off[i], m_off[i + 1])
        for col in ("open", "high", "low", "close", "volume"):
            assert _eq(h(getattr(res.minute, col))[sm], w2[col]), (i, col)
        assert _eq(h(res.minute_ts)[sm], w1["date"][w2["row"]])
        sc = slice(c_off[i], c_off[i + 1])
        assert _eq(h(res.candles.ts)[sc], w3["ts"])
        for col in ("open", "high", "low", "close", "volume"):
            assert _eq(h(getattr(res.candles, col))[sc], w3[col]), (i, col)
        n_ok += 1
    assert n_ok >= 12
    tab = res.candle_table(first_contract=100)
    assert tab["contract"].numel() == res.candles.total and int(tab["contract"].min()) >= 100


def test_random_tape_is_a_function_of_seed_stream_and_index(dev):
    """Element i depends on (seed, stream, i) only: any prefix or odd length gives the same numbers; uniforms in [0, 1) with the
    right moments, exponentials positive with mean 1; different streams differ."""
    import torch
    from iv_interpolation_b200 import device as D
    like = torch.zeros(1, dtype=torch.float64, device=dev)
    a = D.random_tape(1_000_001, 42, 3, "uniform", like=like).cpu().numpy()[:1_000_001]
    b = D.random_tape(4097, 42, 3, "uniform", like=like).cpu().numpy()[:4097]
    assert np.array_equal(a[:4097], b)
    buf = torch.zeros(4099, dtype=torch.float64, device=dev)
    D.random_tape(4097, 42, 3, "uniform", out=buf[1:])                     # 8-byte aligned only: the scalar store path
    assert np.array_equal(buf.cpu().numpy()[1:4098], b) and buf[-1].item() == 0.0
    assert a.min() >= 0.0 and a.max() < 1.0 and abs(a.mean() - 0.5) < 2e-3 and abs(a.var() - 1 / 12) < 1e-3
    c = D.random_tape(1_000_000, 42, 4, "uniform", like=like).cpu().numpy()[:1_000_000]
    assert not np.array_equal(a[:1_000_000], c) and abs(np.corrcoef(a[:1_000_000], c)[0, 1]) < 5e-3
    ex = D.random_tape(1_000_000, 7, 0, "exponential", like=like).cpu().numpy()[:1_000_000]
    assert ex.min() >= 0.0 and abs(ex.mean() - 1.0) < 5e-3 and abs(ex.var() - 1.0) < 2e-2


def test_gather_rows(dev):
    import torch
    from iv_interpolation_b200 import device as D
    src_i = torch.arange(1000, dtype=torch.int64, device=dev) * 7
    src_f = src_i.double() + 0.5
    idx = torch.tensor([0, 999, 5, -1, 1000, 17], dtype=torch.int64, device=dev)
    gi, gf = D.gather(idx, src_i).cpu().numpy(), D.gather(idx, src_f).cpu().numpy()
    assert list(gi) == [0, 6993, 35, np.iinfo(np.int64).min, np.iinfo(np.int64).min, 119]
    assert list(gf[[0, 1, 2, 5]]) == [0.5, 6993.5, 35.5, 119.5] and np.isnan(gf[3]) and np.isnan(gf[4])


def test_host_stream_equals_resident_and_counts_its_copies(dev):
    """HostContractStream (pinned host columns, batches cut by minute rows, candle tables back to pinned host memory) gives the
    candles of one resident ContractPipeline.run over the same contracts — the device tapes are keyed by the global contract id
    of each batch's first contract, so batching at the same cuts reproduces them."""
    import torch
    from iv_interpolation_b200 import contracts as CT, tapes
    tp = tapes.make_underlying_tape(77, 40, rows_per_contract=49)
    shard = CT.HostContractShard.from_tape(tp, ids=500 + np.arange(tp.n_sym))
    stream = CT.HostContractStream(dev, max_rows_per_batch=12 * 2881, seed=9)
    cuts = stream.batches(shard)
    assert len(cuts) == 4 and cuts[0] == (0, 12) and cuts[-1][1] == 40
    out = stream.run(shard)
    assert out["batches"] == 4 and 39 * 2881 < out["interp_rows"] <= 40 * 2881        # dropna cuts the rows before a leading NaN iv
    assert stream.h2d_bytes == (40 + 4) * 8 + tp.n_rows * 80 and stream.d2h_bytes == out["candles"] * 52 + (40 + 4) * 8
    pipe = CT.ContractPipeline(dev, seed=9)
    h = lambda x: x.cpu().numpy()
    n_rows = 0
    for (a, b), tab in zip(cuts, out["tables"]):
        sub = tp.select(np.arange(a, b))
        res = pipe.run(_t(sub.off, dev), _t(sub.date, dev), [_t(sub.cols[c], dev) for c in tapes.INTERP_COLS],
                       [_t(sub.cols[c], dev) for c in tapes.FILL_NUM_COLS], first_contract=500 + a)
        want = res.candle_table(a)
        n_rows += res.interp.total
        for k in want:
            assert _eq(h(tab[k]), h(want[k])), k
    assert n_rows == out["interp_rows"]
    # the last batch is what the pinned ring holds
    torch.cuda.synchronize()
    m = out["tables"][-1]["timestamp"].numel()
    assert _eq(stream.ring["close"][:m].numpy(), h(out["tables"][-1]["close"]))


@pytest.mark.parametrize("strategy", ["spread_simulation", "pipeline", "simple_spread", "price_as_midpoint"])
def test_bridge_with_generated_tapes_equals_the_bridge_on_written_tapes(dev, strategy):
    """ivx_bridge_ohlcv_philox draws inside the kernel what ivx_random_tape would have written: the candles (and the quality bits)
    are bit-equal to the two-step form, per segment and per symbol, with rows that emit nothing, NULL volumes (exponential
    draws) and symbols shorter and longer than a 2048-row segment; the stamps it copies are the gathered ones."""
    import torch
    from iv_interpolation_b200 import _lib, device as D
    rng = np.random.default_rng(8)
    lens = np.array([5, 0, 3000, 1, 2047, 2049, 700, 4500, 33], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)])
    n = int(off[-1])
    base = 100.0 + np.cumsum(rng.normal(0, 0.05, n))
    base[rng.random(n) < 0.03] = np.nan
    base[rng.random(n) < 0.01] = -1.0
    vol = rng.exponential(5.0, n)
    vol[rng.random(n) < 0.2] = np.nan
    vol[rng.random(n) < 0.05] = 0.0
    ts = (np.arange(n, dtype=np.int64) * 60_000_000 + 1_700_000_000_000_000)
    d_off, d_base, d_vol, d_ts = _t(off, dev), _t(base, dev), _t(vol, dev), _t(ts, dev)
    plan = D.bridge_plan(d_off, d_base, d_vol)
    assert 0 < plan.total_e < plan.total < n
    k = _lib.BRIDGE_U_PER_ROW[_lib.BRIDGE_STRATEGY[strategy]]
    u = D.random_tape(k * plan.total, 1234, 6, "uniform", like=d_base)
    e = D.random_tape(plan.total_e, 1234, 7, "exponential", like=d_base)
    want = D.bridge_ohlcv(d_off, d_base, d_vol, plan, strategy, u=u, e=e)
    cols = ("open", "high", "low", "close", "volume", "source_price")
    for seg in (True, False):
        got = D.bridge_ohlcv_philox(d_off, d_base, d_vol, plan, strategy, seed=1234, stream_u=6, stream_e=7, ts=d_ts, segmented=seg)
        assert torch.equal(got.row, want.row) and torch.equal(got.sym_quality, want.sym_quality)
        assert torch.equal(got.ts, d_ts[want.row])
        for c in cols:
            assert torch.equal(getattr(got, c).view(torch.int64), getattr(want, c).view(torch.int64)), (strategy, seg, c)
    bare = D.bridge_ohlcv_philox(d_off, d_base, d_vol, plan, strategy, seed=1234, stream_u=6, stream_e=7, ts=d_ts, want_row=False)
    assert bare.row is None and torch.equal(bare.ts, d_ts[want.row]) and torch.equal(bare.close, want.close)
    other = D.bridge_ohlcv_philox(d_off, d_base, d_vol, plan, strategy, seed=1235, stream_u=6, stream_e=7)
    assert other.ts is None and not torch.equal(other.close, want.close)
    with pytest.raises(_lib.IvxError):
        D.bridge_ohlcv_philox(d_off, d_base, d_vol, plan, "trend_following", seed=1, stream_u=0, stream_e=1)


def test_contract_pipeline_fused_tapes_equal_the_two_step_form(dev):
    from iv_interpolation_b200 import contracts as CT, tapes
    import torch
    tp = tapes.make_underlying_tape(5, 30, rows_per_contract=97)
    args = (_t(tp.off, dev), _t(tp.date, dev), [_t(tp.cols[c], dev) for c in tapes.INTERP_COLS], [_t(tp.cols[c], dev) for c in tapes.FILL_NUM_COLS])
    a = CT.ContractPipeline(dev, seed=3, fused_tapes=True).run(*args, first_contract=40)
    b = CT.ContractPipeline(dev, seed=3, fused_tapes=False).run(*args, first_contract=40)
    assert a.candles.total == b.candles.total > 0 and torch.equal(a.minute_ts, b.minute_ts) and torch.equal(a.minute.row, b.minute.row)
    for c in ("open", "high", "low", "close", "volume"):
        assert torch.equal(getattr(a.minute, c).view(torch.int64), getattr(b.minute, c).view(torch.int64)), c
        assert torch.equal(getattr(a.candles, c).view(torch.int64), getattr(b.candles, c).view(torch.int64)), c
    assert torch.equal(a.candles.ts, b.candles.ts)
