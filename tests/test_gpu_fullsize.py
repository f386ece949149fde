"""Size-independent properties at the bench's (BASELINE-shaped) sizes, where the CPU oracle would take hours:
kernel A, the bridge and kernel B on 2 400 contracts x 20 days = 69 M minute rows; the surface pipeline on 1 200 stamps
x 30 expiries x 2 000 strikes = 72 M quotes.  The small-size bit-exact parity is in the other GPU test files; here the
same code paths run at full width and are held to idempotence, conservation and two-route equalities."""
import numpy as np
import pytest

from iv_interpolation_b200 import tapes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import torch
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _t(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _same(a, b):
    import torch
    return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())


def test_config3_shape_A_bridge_B_properties(dev):
    import torch
    from iv_interpolation_b200 import device as D
    from oracle import np_oracle as O
    tp = tapes.make_contract_tape(3, 2400, rows_per_contract=481)
    kw = dict(step=tapes.MINUTE_US, max_span=30 * tapes.DAY_US)
    off, date = _t(tp.off, dev), _t(tp.date, dev)
    icols = [_t(tp.cols[c], dev) for c in tapes.INTERP_COLS]
    fcols = [_t(tp.cols[c], dev) for c in tapes.FILL_NUM_COLS]
    res = D.interp_linear_ffill(off, date, icols, fcols, [None], **kw)
    n = res.total
    assert n > 68_000_000 and int(res.out_off[-1]) == n
    # ---- A: a sample of contracts bit-equal to the oracle, at their place inside the 69 M-row batch
    o_h = res.out_off.cpu().numpy()
    for i in (0, 1, 777, 1203, 2399):
        a, b = int(tp.off[i]), int(tp.off[i + 1])
        w = O.interp_linear_ffill(tp.date[a:b], [tp.cols[c][a:b] for c in tapes.INTERP_COLS], [tp.cols[c][a:b] for c in tapes.FILL_NUM_COLS],
                                  [np.ones(b - a, bool)], **kw)
        sl = slice(int(o_h[i]), int(o_h[i + 1]))
        assert w["status"] == 0 and sl.stop - sl.start == len(w["date"])
        assert np.array_equal(res.ts[sl].cpu().numpy(), w["date"])
        for k in range(3):
            assert np.array_equal(res.interp[k][sl].cpu().numpy(), w["interp"][k], equal_nan=True), (i, k)
    # ---- A: the minute grid (60 s steps inside a contract), no NaN in the dropna columns, linear interpolation stays
    #      inside the range of its knots
    d = res.ts[1:] - res.ts[:-1]
    starts = torch.zeros(n, dtype=torch.bool, device=dev)
    starts[res.out_off[:-1][res.out_off[:-1] < n]] = True
    assert bool(((d == tapes.MINUTE_US) | starts[1:]).all())
    for k in range(3):
        assert not bool(torch.isnan(res.interp[k]).any())
    iv_src = icols[0][~torch.isnan(icols[0])]
    assert float(res.interp[0].min()) >= float(iv_src.min()) and float(res.interp[0].max()) <= float(iv_src.max())
    # ---- A is idempotent: a complete minute grid interpolates to itself, bit for bit (69 M rows in, 69 M rows out)
    again = D.interp_linear_ffill(res.out_off, res.ts, res.interp, res.fill, [None], **kw)
    assert again.total == n and torch.equal(again.out_off, res.out_off) and torch.equal(again.ts, res.ts)
    for k in range(3):
        assert torch.equal(again.interp[k], res.interp[k])
    for k in range(len(res.fill)):
        assert _same(again.fill[k], res.fill[k])
    del again
    # ---- bridge: one warp per segment == one warp per symbol, bit for bit, on 69 M rows; OHLC ordering; 4-decimal grid
    S, vol = res.interp[1], res.fill[4]
    plan = D.bridge_plan(res.out_off, S, vol)
    assert plan.total == n                                      # every interpolated row has a positive price
    g = torch.Generator(device=dev); g.manual_seed(1)
    u = torch.rand(5 * plan.total, dtype=torch.float64, device=dev, generator=g)
    e = torch.empty(max(1, plan.total_e), dtype=torch.float64, device=dev).exponential_(generator=g)
    seg = D.bridge_ohlcv(res.out_off, S, vol, plan, "spread_simulation", u=u, e=e, segmented=True)
    sym = D.bridge_ohlcv(res.out_off, S, vol, plan, "spread_simulation", u=u, e=e, segmented=False)
    for c in ("open", "high", "low", "close", "volume", "source_price"):
        assert torch.equal(getattr(seg, c), getattr(sym, c)), c
    assert torch.equal(seg.row, sym.row) and torch.equal(seg.sym_quality, sym.sym_quality) and not bool(seg.sym_quality.any())
    del sym, u, e
    assert bool((seg.low <= torch.minimum(seg.open, seg.close)).all()) and bool((seg.high >= torch.maximum(seg.open, seg.close)).all())
    # closes sit on numpy's 4-decimal grid (a tensor divisor: torch turns division by a Python scalar into a multiplication)
    ten4 = torch.tensor(1e4, dtype=torch.float64, device=dev)
    assert bool((torch.round(seg.close * 1e4) / ten4 == seg.close).all()) and torch.equal(seg.source_price, S)
    assert bool((seg.volume >= 0).all()) and int(torch.isnan(seg.volume).sum()) == 0
    # ---- B on the bridge's candles: complete 5-minute buckets only, extrema and volume conserved
    cb = D.ohlcv_reduce(res.out_off, res.ts, seg.open, seg.high, seg.low, seg.close, seg.volume, bucket_ticks=5 * tapes.MINUTE_US, min_count=5)
    m = cb.total
    assert 13_000_000 < m <= n // 5 and bool((cb.count[:m] == 5).all())          # unique 1-minute stamps: a complete bucket has 5 rows
    assert bool((cb.high[:m] >= cb.low[:m]).all())
    assert float(cb.high[:m].max()) <= float(seg.high.max()) and float(cb.low[:m].min()) >= float(seg.low.min())
    # the rows of complete buckets: bucket id changes every 5 rows; volume is conserved over them (Kahan vs pairwise: 1e-12)
    bucket = torch.div(res.ts, 5 * tapes.MINUTE_US, rounding_mode="floor")
    symid = torch.searchsorted(res.out_off[1:].contiguous(), torch.arange(n, device=dev), right=True)
    key = symid * (1 << 40) + (bucket - bucket.min())
    _, inv, cnt = torch.unique_consecutive(key, return_inverse=True, return_counts=True)
    kept_rows = cnt[inv] >= 5
    assert int(kept_rows.sum()) == 5 * m
    tot_in = float(seg.volume[kept_rows].sum())
    assert abs(float(cb.volume[:m].sum()) - tot_in) <= 1e-10 * tot_in


def test_config2_shape_surface_two_routes(dev):
    """72 M quotes: the per-quote layout (general spline) and the chain layout (shared ladders, precomputed logs) are two
    independent routes through different kernels; their IVs, surfaces and candles must agree to the IV tolerance."""
    import torch
    from iv_interpolation_b200 import surface as SF
    spec = SF.ChainSpec(1200, 30, 2000, 512, 64, bucket_ticks=60 * tapes.US, min_count=10)
    chain = tapes.make_quote_chain(9, spec, dev)
    a = SF.SurfacePipeline(spec, dev).run(chain)
    iv_a, close_a, nb = a.iv.clone(), a.close.clone(), a.n_buckets
    ok = ~torch.isnan(iv_a)
    assert float(ok.double().mean()) > 0.99
    err = (iv_a[ok] - chain.sigma_true[ok]).abs()                                 # round trip sigma -> price -> sigma at full width
    # every quote within 1e-8 plus what its own conditioning allows (64 ulps of the inverted quantity over vega): no quota, no 1e-5
    from tests.ivcond import iv_allowed
    g = lambda name: getattr(chain, name)[ok]
    allowed = iv_allowed(g("price"), g("S"), g("K"), g("T"), g("r"), g("is_call"), chain.sigma_true[ok], xp=torch)
    assert bool((err <= allowed).all()), f"worst excess {float((err - allowed).max()):.3e}"
    assert float((allowed <= 2e-8).double().mean()) >= 0.99 and float((err <= 1e-8).double().mean()) >= 0.999
    del a
    sc = SF.StructuredChain.from_quotes(chain)
    b = SF.SurfacePipeline(spec, dev).run(sc)
    assert b.n_buckets == nb == 20
    assert bool((torch.isnan(b.iv) == torch.isnan(iv_a)).all()) and float((b.iv[ok] - iv_a[ok]).abs().max()) <= 1e-9
    ca, cb = close_a[:nb * spec.n_series], b.close[:nb * spec.n_series]
    fin = ~torch.isnan(ca)
    assert bool((torch.isnan(cb) == ~fin).all()) and float((ca[fin] - cb[fin]).abs().max()) <= 1e-8


def test_config2_shape_gappy_tape_filled_band_equals_exact_solve(dev):
    """A full-width chain (600 stamps x 30 expiries x 2000 strikes) with 0.1 % of its quotes missing — 86 % of its slices have a gap.
    The banded strike spline fills isolated gaps on the tensor path and hands the rest to the exact solve; the scan kernel solves every
    slice exactly on its quoted knots.  The two smiles agree to 1e-9 everywhere, the band keeps the great majority of the slices,
    and the candles that come out are the same."""
    import torch
    from iv_interpolation_b200 import surface as SF
    spec = SF.ChainSpec(600, 30, 2000, 512, 64, bucket_ticks=60 * tapes.US, min_count=30)
    chain = tapes.make_quote_chain(77, spec, dev, drop_frac=0.001)
    gaps = torch.isnan(chain.price).view(-1, spec.n_strike).any(dim=1).double().mean().item()
    assert 0.7 < gaps < 0.95
    band = SF.SurfacePipeline(spec, dev, banded=True)
    rb = band.run(chain)
    smile_b, close_b = rb.smile.clone(), rb.close.clone()
    flagged = int(band._fb[1].item())
    assert flagged < 0.10 * spec.n_t * spec.n_exp, flagged        # gaps near an end or next to each other only
    del band, rb
    torch.cuda.empty_cache()
    scan = SF.SurfacePipeline(spec, dev, banded=False)
    rs = scan.run(chain)
    nan_b, nan_s = torch.isnan(smile_b), torch.isnan(rs.smile)
    assert bool((nan_b == nan_s).all())
    err = (smile_b - rs.smile)[~nan_b].abs().max().item()
    assert err <= 1e-9, err
    cb, cs = close_b, rs.close
    assert bool((torch.isnan(cb) == torch.isnan(cs)).all()) and (cb - cs)[~torch.isnan(cb)].abs().max().item() <= 1e-8
