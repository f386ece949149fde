"""GPU parity of the north_star surface pipeline (kernels C -> D -> D' -> B'), stage by stage against the
scipy/numpy oracle: IV and spline values within 1e-8 absolute (BASELINE.json north_star), candle
selection bit-exact.  The reference has no code for these stages ("parity unpinned")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from iv_interpolation_b200 import tapes  # noqa: E402
from oracle import np_oracle as O, surface_oracle as SO  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from iv_interpolation_b200 import _lib
    _lib.lib()
    return torch.device("cuda")


def _eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())


def _chain(dev, seed=5, n_t=130, n_exp=5, n_strike=64, nk=48, nm=12, drop=0.02, min_count=30):
    from iv_interpolation_b200 import surface as SF
    spec = SF.ChainSpec(n_t, n_exp, n_strike, nk, nm, bucket_ticks=60 * tapes.US, min_count=min_count)
    return spec, tapes.make_quote_chain(seed, spec, dev, drop_frac=drop)


def _stagewise_check(spec, chain, res):
    h = lambda t: t.cpu().numpy()
    price, S, K, T, r, call = (h(getattr(chain, k)) for k in ("price", "S", "K", "T", "r", "is_call"))
    n_t, n_e, n_k = spec.n_t, spec.n_exp, spec.n_strike
    iv = h(res.iv)
    want = SO.implied_vol_fast(price, S, K, T, r, call)
    with np.errstate(all="ignore"):
        kd = K * np.exp(-r * T)
        tv = price - np.maximum(np.where(call != 0, S - kd, kd - S), 0.0)
    well = ~np.isnan(want) & (tv > 1e-6 * S)
    assert well.mean() > 0.85
    assert not np.isnan(iv[well]).any()
    assert np.abs(iv - want)[well].max() <= 1e-8, "implied vol"
    assert np.isnan(iv[np.isnan(price)]).all()
    # later stages take the device's own IVs so that each stage is checked on identical inputs
    smile = h(res.smile).reshape(n_t, n_e, spec.n_grid_k)
    w_smile = SO.smile_stage(K, iv, n_t, n_e, n_k, h(chain.grid_k), spec.extrap)
    assert _eq(np.isnan(smile), np.isnan(w_smile))
    assert np.nanmax(np.abs(smile - w_smile)) <= 1e-8, "strike spline"
    surface = h(res.surface).reshape(n_t, spec.n_grid_m, spec.n_grid_k)
    w_surf = SO.maturity_stage(T, smile, n_t, n_e, n_k, h(chain.grid_m), spec.extrap)
    assert _eq(np.isnan(surface), np.isnan(w_surf))
    assert np.nanmax(np.abs(surface - w_surf)) <= 1e-8, "maturity spline"
    w = SO.candle_stage(h(chain.ts), surface, spec.bucket_ticks, spec.min_count)
    assert res.n_buckets == len(w["count"]) and _eq(h(res.bucket_ts), w["bucket_ts"]) and _eq(h(res.count), w["count"])
    for k in ("open", "high", "low", "close"):
        assert _eq(h(getattr(res, k)).reshape(res.n_buckets, -1), w[k]), f"candles: {k}"


def test_surface_pipeline_stagewise(dev):
    from iv_interpolation_b200 import surface as SF
    spec, chain = _chain(dev)
    pipe = SF.SurfacePipeline(spec, dev)
    res = pipe.run(chain)
    _stagewise_check(spec, chain, res)
    # a second pass over the same buffers gives the same bits (no state leaks between steps)
    first = [t.clone() for t in (res.iv, res.surface, res.close)]
    res2 = pipe.run(chain)
    for a, b in zip(first, (res2.iv, res2.surface, res2.close)):
        assert _eq(a.cpu().numpy(), b.cpu().numpy())


def test_chain_layout_matches_per_quote_layout(dev):
    """The chain layout (S per stamp, K per contract, T per (stamp, expiry): 9 B/quote) passes the same stage-wise
    oracle checks and agrees with the per-quote layout to far below the tolerance."""
    from iv_interpolation_b200 import surface as SF
    spec, chain = _chain(dev, seed=6, n_t=70, n_exp=6, n_strike=300, nk=64, nm=16, drop=0.03, min_count=10)
    sc = SF.StructuredChain.from_quotes(chain)
    assert sc.S_t.numel() == 70 and sc.K_ek.numel() == 1800 and sc.T_te.numel() == 420 and sc.is_call.numel() == 1800
    a = SF.SurfacePipeline(spec, dev).run(chain)
    b = SF.SurfacePipeline(spec, dev).run(sc)
    # call/put flags given per quote instead of per contract: same bits
    import dataclasses
    b2 = SF.SurfacePipeline(spec, dev).run(dataclasses.replace(sc, is_call=chain.is_call))
    assert _eq(b.iv.cpu().numpy(), b2.iv.cpu().numpy())
    _stagewise_check(spec, chain, b)
    iva, ivb = a.iv.cpu().numpy(), b.iv.cpu().numpy()
    assert _eq(np.isnan(iva), np.isnan(ivb))
    price, S, K, T, r, call = (getattr(chain, k).cpu().numpy() for k in ("price", "S", "K", "T", "r", "is_call"))
    with np.errstate(all="ignore"):
        kd = K * np.exp(-r * T)
        tv = price - np.maximum(np.where(call != 0, S - kd, kd - S), 0.0)
    well = ~np.isnan(iva) & (tv > 1e-6 * S)
    assert np.abs(iva - ivb)[well].max() <= 1e-11
    assert np.nanmax(np.abs(a.surface.cpu().numpy() - b.surface.cpu().numpy())) <= 1e-8


def test_surface_pipeline_full_width_slices(dev):
    """2000-strike slices and the 512 x 64 grid of BASELINE.json configs[1]/[3], a few stamps."""
    from iv_interpolation_b200 import surface as SF
    spec, chain = _chain(dev, seed=9, n_t=6, n_exp=30, n_strike=2000, nk=512, nm=64, drop=0.01, min_count=1)
    res = SF.SurfacePipeline(spec, dev).run(chain)
    _stagewise_check(spec, chain, res)


def test_surface_pipeline_moving_ladders(dev):
    """Per-quote layout at full width when the strike ladders are NOT the same at every stamp: the spline presumes the first
    stamp's ladders and verifies per slice; stamps whose strikes moved (even by one ulp in one knot) must take the general
    solve and still pass the oracle; a rewritten K column refactors the tables."""
    from iv_interpolation_b200 import device as D, surface as SF
    spec, chain = _chain(dev, seed=12, n_t=8, n_exp=30, n_strike=2000, nk=512, nm=64, drop=0.0, min_count=1)
    per_t = spec.n_exp * spec.n_strike
    K = chain.K.clone()
    K[2 * per_t:3 * per_t] *= 1.001                                            # a whole stamp on other strikes
    K[5 * per_t + 777] = torch.nextafter(K[5 * per_t + 777], K.new_tensor(float("inf")))   # one knot, one ulp
    chain.K = K
    chain.price = D.bs_price(chain.S, chain.K, chain.T, chain.r, chain.sigma_true, chain.is_call)
    pipe = SF.SurfacePipeline(spec, dev)
    res = pipe.run(chain)
    _stagewise_check(spec, chain, res)
    # the same pipeline object, the K column rewritten in place: stale tables would be wrong for the slices that now match
    first_smile = res.smile.clone()
    chain.K.mul_(1.0005)
    chain.price.copy_(D.bs_price(chain.S, chain.K, chain.T, chain.r, chain.sigma_true, chain.is_call))
    res2 = pipe.run(chain)
    _stagewise_check(spec, chain, res2)
    assert not _eq(first_smile.cpu().numpy(), res2.smile.cpu().numpy())


def test_host_stream_equals_resident_pipeline(dev):
    """The host-facing entry (pinned host columns in, host candles out, double-buffered chunks) gives the same
    bits as the resident pipeline over the whole tape."""
    from iv_interpolation_b200 import surface as SF
    chunk_t, n_chunks = 120, 5
    spec, chain = _chain(dev, seed=11, n_t=chunk_t * n_chunks, n_exp=4, n_strike=48, nk=32, nm=8, drop=0.05, min_count=1)
    full = SF.SurfacePipeline(spec, dev).run(chain)
    cspec = SF.ChainSpec(chunk_t, 4, 48, 32, 8, bucket_ticks=spec.bucket_ticks, min_count=1)
    per = cspec.n_quotes
    chunks = []
    for i in range(n_chunks):
        sl = slice(i * per, (i + 1) * per)
        cols = {k: getattr(chain, k)[sl].cpu().pin_memory() for k in SF.QUOTE_F64 + ("is_call",)}
        chunks.append(SF.HostQuoteChunk(ts=chain.ts[i * chunk_t:(i + 1) * chunk_t].cpu().pin_memory(), **cols))
    stream = SF.HostSurfaceStream(cspec, dev, chain.grid_k, chain.grid_m, max_buckets_per_chunk=3)
    nb = n_chunks * 3
    out = [torch.full((nb * cspec.n_series,), -1.0, dtype=torch.float64).pin_memory() for _ in range(4)]
    out_ts = torch.zeros(nb, dtype=torch.int64).pin_memory()
    out_cnt = torch.zeros(nb, dtype=torch.int32).pin_memory()
    rows = stream.run(chunks, out, out_ts, out_cnt)
    assert rows == full.n_buckets == 10
    assert _eq(out_ts[:rows].numpy(), full.bucket_ts.cpu().numpy()) and _eq(out_cnt[:rows].numpy(), full.count.cpu().numpy())
    for got, want in zip(out, (full.open, full.high, full.low, full.close)):
        assert _eq(got[: rows * cspec.n_series].numpy(), want.cpu().numpy())
    assert stream.h2d_bytes == n_chunks * (per * 41 + chunk_t * 8)
    # the same tape in the chain layout: 8 B per quote over PCIe, same candles to within the IV tolerance
    sc = SF.StructuredChain.from_quotes(chain)
    cchunks = []
    for i in range(n_chunks):
        sl, tl = slice(i * per, (i + 1) * per), slice(i * chunk_t, (i + 1) * chunk_t)
        pin = lambda t: t.cpu().pin_memory()
        cchunks.append(SF.HostChainChunk(ts=pin(sc.ts[tl]), price=pin(sc.price[sl]), S_t=pin(sc.S_t[tl]),
                                         T_te=pin(sc.T_te[i * chunk_t * 4:(i + 1) * chunk_t * 4]), r_t=pin(sc.r_t[tl])))
    cstream = SF.HostSurfaceStream(cspec, dev, chain.grid_k, chain.grid_m, max_buckets_per_chunk=3, K_ek=sc.K_ek, is_call_ek=sc.is_call)
    out2 = [torch.full((nb * cspec.n_series,), -1.0, dtype=torch.float64).pin_memory() for _ in range(4)]
    rows2 = cstream.run(cchunks, out2, out_ts, out_cnt, bucket_counts=[2] * n_chunks)
    assert rows2 == rows and cstream.h2d_bytes == n_chunks * (per * 8 + chunk_t * (8 + 8 + 8 + 4 * 8)) and sc.is_call.numel() == 4 * 48
    for got, want in zip(out2, out):
        g, w = got[: rows * cspec.n_series].numpy(), want[: rows * cspec.n_series].numpy()
        assert _eq(np.isnan(g), np.isnan(w)) and np.nanmax(np.abs(g - w)) <= 1e-8
    with pytest.raises(Exception):
        cstream.run(chunks[:1], out2, out_ts, out_cnt)        # wrong chunk layout for this stream


def test_tick_price_layout_equals_prices_rounded_to_four_decimals(dev):
    """The 4-byte wire format (int32 ticks of 1e-4, the reference's DECIMAL(15,4); INT32_MIN = no quote): the chain solver on
    ticks is bit-equal to the chain solver on round(price, 4), a missing quote stays missing, and the host stream in the tick
    layout returns the candles of the chain layout fed with the rounded prices."""
    import torch
    from iv_interpolation_b200 import device as D, surface as SF
    spec = SF.ChainSpec(120, 4, 96, 32, 8, bucket_ticks=60 * tapes.US, min_count=10)
    chain = tapes.make_quote_chain(17, spec, dev, drop_frac=0.01)
    sc = SF.StructuredChain.from_quotes(chain)
    ticks = SF.HostTickChainChunk.encode(sc.price)
    # round(price, 4) on the host (IEEE division; torch's CUDA division by a scalar multiplies by the reciprocal)
    p_host = sc.price.cpu().numpy()
    rounded = torch.from_numpy(np.where(np.isnan(p_host), p_host, np.rint(p_host * 1e4) / 1e4)).to(dev)
    assert bool((ticks[torch.isnan(sc.price)] == -2 ** 31).all()) and int(torch.isnan(sc.price).sum()) > 10
    kw = dict(n_t=spec.n_t, n_exp=spec.n_exp, n_strike=spec.n_strike)
    a = D.bs_implied_vol_chain(rounded, sc.is_call, sc.S_t, sc.K_ek, sc.T_te, sc.r_t, **kw)
    b = D.bs_implied_vol_chain(ticks, sc.is_call, sc.S_t, sc.K_ek, sc.T_te, sc.r_t, **kw)
    assert bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
    assert bool(torch.isnan(b[torch.isnan(sc.price)]).all()) and float(torch.isnan(b).double().mean()) < 0.1
    # host stream, two chunks of 60 stamps
    cspec = SF.ChainSpec(60, 4, 96, 32, 8, bucket_ticks=60 * tapes.US, min_count=10)
    pin = lambda t: t.cpu().pin_memory()
    def chunks(price, cls):
        out = []
        for i in range(2):
            tl, sl = slice(i * 60, (i + 1) * 60), slice(i * cspec.n_quotes, (i + 1) * cspec.n_quotes)
            out.append(cls(ts=pin(sc.ts[tl]), price=pin(price[sl]), S_t=pin(sc.S_t[tl]), T_te=pin(sc.T_te[i * 60 * 4:(i + 1) * 60 * 4]), r_t=pin(sc.r_t[tl])))
        return out
    def run(price, cls, ticks_):
        st = SF.HostSurfaceStream(cspec, dev, chain.grid_k, chain.grid_m, max_buckets_per_chunk=2, K_ek=sc.K_ek, is_call_ek=sc.is_call, price_ticks=ticks_)
        out = [torch.empty(4 * cspec.n_series, dtype=torch.float64).pin_memory() for _ in range(4)]
        ts, cnt = torch.empty(4, dtype=torch.int64).pin_memory(), torch.empty(4, dtype=torch.int32).pin_memory()
        nb = st.run(chunks(price, cls), out, ts, cnt)
        return nb, [o[: nb * cspec.n_series].clone() for o in out], st.h2d_bytes
    nb1, c1, bytes1 = run(rounded, SF.HostChainChunk, False)
    nb2, c2, bytes2 = run(ticks, SF.HostTickChainChunk, True)
    assert nb1 == nb2 == 2 and bytes2 < 0.55 * bytes1
    for x, y in zip(c1, c2):
        assert bool(((x == y) | (torch.isnan(x) & torch.isnan(y))).all())
    with pytest.raises(Exception):
        SF.HostSurfaceStream(cspec, dev, chain.grid_k, chain.grid_m, max_buckets_per_chunk=2, K_ek=sc.K_ek, is_call_ek=sc.is_call).run(
            chunks(ticks, SF.HostTickChainChunk), c1, torch.empty(4, dtype=torch.int64), torch.empty(4, dtype=torch.int32))
