"""Seeded synthetic quote tapes for the BASELINE.json configs.

All randomness is ``np.random.default_rng(seed)``; stamps are unique per contract and tz-naive.
The generators return structure-of-arrays tapes (CSR-packed by contract) and can render any
contract as the DataFrame the reference's classes take.  The value distributions mimic the
reference's own sample generator (reference main.py:185-226: random-walk 1-min candles,
``volume ~ Exp``) — no reference code is reused.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

US = 1_000_000            # ticks per second at datetime64[us], pandas 3's inferred unit
MINUTE_US = 60 * US
HOUR_US = 3600 * US
DAY_US = 86400 * US
T0_US = 1_678_838_400 * US   # 2023-03-15 00:00:00 UTC, inside the reference's own window (batch_processor.py:30-33)

INTERP_COLS = ("iv", "underlying_price", "time_to_maturity")
FILL_NUM_COLS = ("strike", "interest_rate", "mark_price", "index_price", "volume", "quote_volume")


@dataclass
class ContractTape:
    """Task-1 input, CSR-packed: rows of contract i are ``off[i]:off[i+1]``, sorted by date."""
    off: np.ndarray                       # int64[n_sym+1]
    date: np.ndarray                      # int64[N] ticks (us)
    cols: Dict[str, np.ndarray]           # float64[N] for INTERP_COLS + FILL_NUM_COLS
    symbols: List[str]                    # contract names (underlying-expiry-strike-c/p, main.py:177-183)
    callput: np.ndarray                   # uint8[n_sym]: 1 call / 0 put
    unit: str = "us"

    @property
    def n_sym(self) -> int:
        return len(self.off) - 1

    @property
    def n_rows(self) -> int:
        return int(self.off[-1])

    def frame(self, i: int):
        """Contract ``i`` as the frame reference batch_processor.py:76-83 would fetch."""
        import pandas as pd

        a, b = int(self.off[i]), int(self.off[i + 1])
        d = {"symbol": self.symbols[i], "date": self.date[a:b].view("datetime64[us]")}
        for c in INTERP_COLS:
            d[c] = self.cols[c][a:b]
        d["strike"] = self.cols["strike"][a:b]
        d["callput"] = "c" if self.callput[i] else "p"
        for c in ("interest_rate", "mark_price", "index_price", "volume", "quote_volume"):
            d[c] = self.cols[c][a:b]
        d["record_time"] = [f"rt{j}" for j in range(a, b)]
        return pd.DataFrame(d)

    def select(self, idx) -> "ContractTape":
        """Sub-tape with the given contracts (used for sharding and bounded CPU samples)."""
        idx = np.asarray(idx, dtype=np.int64)
        lens = self.off[idx + 1] - self.off[idx]
        off = np.zeros(len(idx) + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        rows = np.concatenate([np.arange(self.off[i], self.off[i + 1]) for i in idx]) if len(idx) else np.zeros(0, np.int64)
        return ContractTape(off, self.date[rows], {k: v[rows] for k, v in self.cols.items()},
                            [self.symbols[i] for i in idx], self.callput[idx], self.unit)


def make_contract_tape(
    seed: int,
    n_contracts: int,
    rows_per_contract: int = 25,
    spacing_us: int = HOUR_US,
    nan_frac: float = 0.05,
    offgrid_frac: float = 0.02,
    missing_frac: float = 0.0,
    underlying: str = "btc",
    t0: int = T0_US,
    s0: float = 25_000.0,
) -> ContractTape:
    """Hourly (or ``spacing_us``) source rows per contract.

    5 % of iv are NaN (some leading/trailing), 2 % of rows sit 30 s off the minute grid,
    ``missing_frac`` of rows are absent altogether (config 5).  SURVEY.md §8(d) config 1.
    """
    rng = np.random.default_rng(seed)
    n_full = rows_per_contract
    offs = [0]
    date, cols = [], {c: [] for c in INTERP_COLS + FILL_NUM_COLS}
    symbols, callput = [], np.zeros(n_contracts, np.uint8)
    base_t = t0 + np.arange(n_full, dtype=np.int64) * spacing_us
    spot = s0 + np.cumsum(rng.normal(0.0, 50.0, n_full))
    strikes = np.round(s0 * np.linspace(0.5, 1.5, max(2, (n_contracts + 11) // 12)) / 100.0) * 100.0
    for i in range(n_contracts):
        keep = np.ones(n_full, bool)
        if missing_frac > 0:
            keep = rng.random(n_full) >= missing_frac
            keep[[0, -1]] = True
        t = base_t[keep].copy()
        m = len(t)
        og = rng.random(m) < offgrid_frac
        og[[0, -1]] = False                         # keep the span ends on the grid
        t[og] += 30 * US
        iv = rng.uniform(0.2, 1.2, m)
        nanm = rng.random(m) < nan_frac
        if i % 7 == 3:
            nanm[0] = True                          # some leading NaN (rows dropped by dropna)
        if i % 11 == 5:
            nanm[-1] = True                         # some trailing NaN (flat-filled)
        iv[nanm] = np.nan
        exp_i, k_i, cp = (i // 2) % 6, (i // 12) % len(strikes), i % 2
        ttm0 = (7 + 30 * exp_i) / 365.0
        date.append(t)
        cols["iv"].append(iv)
        cols["underlying_price"].append(spot[keep] + rng.normal(0, 1.0, m))
        cols["time_to_maturity"].append(ttm0 - (t - t0) / (365.0 * DAY_US))
        cols["strike"].append(np.full(m, strikes[k_i]))
        cols["interest_rate"].append(np.full(m, 0.05))
        cols["mark_price"].append(rng.uniform(0.001, 0.2, m))
        cols["index_price"].append(spot[keep])
        vol = rng.exponential(5.0, m)
        vol[rng.random(m) < 0.03] = np.nan          # NULL volumes are forward-filled column-wise
        cols["volume"].append(vol)
        cols["quote_volume"].append(vol * spot[keep])
        callput[i] = cp
        symbols.append(f"{underlying}-{15 + exp_i}mar23-{int(strikes[k_i])}-{'c' if cp else 'p'}")
        offs.append(offs[-1] + m)
    return ContractTape(
        np.asarray(offs, np.int64), np.concatenate(date),
        {k: np.concatenate(v) for k, v in cols.items()}, symbols, callput,
    )


def make_underlying_tape(seed: int, n_contracts: int, rows_per_contract: int = 481, spacing_us: int = HOUR_US,
                         nan_frac: float = 0.05, offgrid_frac: float = 0.02, underlying: str = "u0", t0: int = T0_US,
                         s0: float = 25_000.0, drop_frac: float = 0.0) -> ContractTape:
    """All contracts of ONE underlying over ``rows_per_contract`` hourly stamps (481 = 20 days): the unit BASELINE.json
    configs[2] shards by.  Same distributions as ``make_contract_tape`` (5 % NaN iv incl. leading/trailing ones, 2 % of
    rows 30 s off the grid, NaN volumes), drawn in whole-underlying arrays so that 100 underlyings take seconds.
    ``drop_frac`` > 0: that share of the hourly rows is missing altogether (the sparse replay tape of configs[4])."""
    rng = np.random.default_rng(seed)
    n, m = int(n_contracts), int(rows_per_contract)
    base_t = t0 + np.arange(m, dtype=np.int64) * spacing_us
    spot = s0 + np.cumsum(rng.normal(0.0, 50.0, m))
    t = np.broadcast_to(base_t, (n, m)).copy()
    og = rng.random((n, m)) < offgrid_frac
    og[:, 0] = og[:, -1] = False
    t[og] += 30 * US
    iv = rng.uniform(0.2, 1.2, (n, m))
    nanm = rng.random((n, m)) < nan_frac
    idx = np.arange(n)
    nanm[idx % 7 == 3, 0] = True
    nanm[idx % 11 == 5, -1] = True
    iv[nanm] = np.nan
    n_strikes = max(2, (n + 11) // 12)
    strikes = np.round(s0 * np.linspace(0.5, 1.5, n_strikes) / 100.0) * 100.0
    exp_i, k_i, cp = (idx // 2) % 6, (idx // 12) % n_strikes, idx % 2
    ttm0 = (7 + 30 * exp_i) / 365.0
    vol = rng.exponential(5.0, (n, m))
    vol[rng.random((n, m)) < 0.03] = np.nan
    cols = {
        "iv": iv,
        "underlying_price": spot[None, :] + rng.normal(0.0, 1.0, (n, m)),
        "time_to_maturity": ttm0[:, None] - (t - t0) / (365.0 * DAY_US),
        "strike": np.broadcast_to(strikes[k_i][:, None], (n, m)),
        "interest_rate": np.full((n, m), 0.05),
        "mark_price": rng.uniform(0.001, 0.2, (n, m)),
        "index_price": np.broadcast_to(spot[None, :], (n, m)),
        "volume": vol,
        "quote_volume": vol * spot[None, :],
    }
    symbols = [f"{underlying}-{15 + int(e)}mar23-{int(strikes[k])}-{'c' if c else 'p'}" for e, k, c in zip(exp_i, k_i, cp)]
    if drop_frac > 0.0:
        # sparse tape (BASELINE.json configs[4]): hourly rows missing at random, never a contract's first or last — ragged CSR
        keep = rng.random((n, m)) >= drop_frac
        keep[:, 0] = keep[:, -1] = True
        off = np.concatenate([[0], np.cumsum(keep.sum(axis=1))]).astype(np.int64)
        return ContractTape(off, t[keep], {k: np.ascontiguousarray(np.asarray(v, dtype=np.float64)[keep]) for k, v in cols.items()},
                            symbols, cp.astype(np.uint8))
    return ContractTape(np.arange(n + 1, dtype=np.int64) * m, t.reshape(-1),
                        {k: np.ascontiguousarray(v, dtype=np.float64).reshape(-1) for k, v in cols.items()}, symbols,
                        cp.astype(np.uint8))


@dataclass
class CandleTape:
    """Task-2 input, CSR-packed 1-min (or finer) OHLCV rows per symbol, sorted by timestamp."""
    off: np.ndarray
    ts: np.ndarray                        # int64[N] ticks (us)
    o: np.ndarray
    h: np.ndarray
    l: np.ndarray
    c: np.ndarray
    v: np.ndarray
    symbols: List[str]
    unit: str = "us"

    @property
    def n_sym(self) -> int:
        return len(self.off) - 1

    @property
    def n_rows(self) -> int:
        return int(self.off[-1])

    def frame(self, i: int):
        """Symbol ``i`` as the frame candle core.py:215-223 would fetch."""
        import pandas as pd

        a, b = int(self.off[i]), int(self.off[i + 1])
        return pd.DataFrame({
            "symbol": self.symbols[i], "timestamp": self.ts[a:b].view("datetime64[us]"),
            "open": self.o[a:b], "high": self.h[a:b], "low": self.l[a:b],
            "close": self.c[a:b], "volume": self.v[a:b],
        })

    def select(self, idx) -> "CandleTape":
        idx = np.asarray(idx, dtype=np.int64)
        lens = self.off[idx + 1] - self.off[idx]
        off = np.zeros(len(idx) + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        rows = np.concatenate([np.arange(self.off[i], self.off[i + 1]) for i in idx]) if len(idx) else np.zeros(0, np.int64)
        return CandleTape(off, self.ts[rows], self.o[rows], self.h[rows], self.l[rows],
                          self.c[rows], self.v[rows], [self.symbols[i] for i in idx], self.unit)


def make_candle_tape(
    seed: int,
    n_symbols: int,
    rows_per_symbol: int = 1441,
    spacing_us: int = MINUTE_US,
    gap_frac: float = 0.01,
    nan_frac: float = 0.01,
    t0: int = T0_US,
    jitter_start: bool = True,
    s0: float = 25_000.0,
) -> CandleTape:
    """Random-walk candles: open=close=S, high=S+|N(0,3)|, low=S-|N(0,3)|, volume~Exp(50)
    (the shape of reference main.py:199-210), ``gap_frac`` of rows missing, ``nan_frac`` NaNs."""
    rng = np.random.default_rng(seed)
    offs = [0]
    ts, O, H, L, C, V = [], [], [], [], [], []
    symbols = []
    for i in range(n_symbols):
        start = t0 + (int(rng.integers(0, 7)) * spacing_us if jitter_start else 0)
        keep = rng.random(rows_per_symbol) >= gap_frac
        t = (start + np.arange(rows_per_symbol, dtype=np.int64) * spacing_us)[keep]
        m = len(t)
        s = s0 + np.cumsum(rng.normal(0, 5.0, m))
        o = s.copy()
        c = s + rng.normal(0, 1.0, m)
        h = np.maximum(o, c) + np.abs(rng.normal(0, 3.0, m))
        l = np.minimum(o, c) - np.abs(rng.normal(0, 3.0, m))
        v = rng.exponential(50.0, m)
        for a in (o, h, l, c, v):
            a[rng.random(m) < nan_frac] = np.nan
        ts.append(t); O.append(o); H.append(h); L.append(l); C.append(c); V.append(v)
        symbols.append(f"btc-sym{i}")
        offs.append(offs[-1] + m)
    cat = np.concatenate
    return CandleTape(np.asarray(offs, np.int64), cat(ts), cat(O), cat(H), cat(L), cat(C), cat(V), symbols)


@dataclass
class QuoteTape:
    """Per-quote SoA for the implied-vol solver: price, S, K, T, r, is_call (49 B/quote with the output)."""
    price: np.ndarray
    S: np.ndarray
    K: np.ndarray
    T: np.ndarray
    r: np.ndarray
    is_call: np.ndarray           # uint8
    sigma_true: Optional[np.ndarray] = None


def make_quote_tape(seed: int, n: int, s0: float = 4500.0, pricer=None) -> QuoteTape:
    """SPX-like quotes: sigma~U(0.05,1.5), K in [0.5S,1.5S], T in [1/365,2], r=0.05.

    ``pricer(S,K,T,r,sigma,is_call)`` supplies the Black-Scholes prices (tests pass the oracle's,
    the bench passes the device generator); without one, prices are left NaN.
    """
    rng = np.random.default_rng(seed)
    S = s0 * (1.0 + 0.01 * rng.standard_normal(n))
    K = S * rng.uniform(0.5, 1.5, n)
    T = rng.uniform(1.0 / 365.0, 2.0, n)
    r = np.full(n, 0.05)
    sigma = rng.uniform(0.05, 1.5, n)
    is_call = (rng.random(n) < 0.5).astype(np.uint8)
    price = np.full(n, np.nan) if pricer is None else np.asarray(pricer(S, K, T, r, sigma, is_call), dtype=np.float64)
    return QuoteTape(price, S, K, T, r, is_call, sigma)


def make_smile_slices(seed: int, n_slices: int, n_knots_lo: int = 50, n_knots_hi: int = 2000,
                      n_query: int = 512, nan_frac: float = 0.0):
    """Config 4: per slice, n_k in [lo,hi] strictly increasing strikes with a smile-shaped IV,
    and one shared n_query-point strike grid.  Returns (knot_off, x, y, xq)."""
    rng = np.random.default_rng(seed)
    nk = rng.integers(n_knots_lo, n_knots_hi + 1, n_slices)
    off = np.zeros(n_slices + 1, np.int64)
    np.cumsum(nk, out=off[1:])
    x = np.empty(int(off[-1]))
    y = np.empty(int(off[-1]))
    for i in range(n_slices):
        gaps = rng.uniform(0.2, 1.8, nk[i])
        xi = 0.5 + np.cumsum(gaps) / gaps.sum()              # moneyness in (0.5, 1.5]
        m = np.log(xi)
        yi = 0.2 + 0.35 * m * m - 0.08 * m + 0.003 * rng.standard_normal(nk[i])
        x[off[i]:off[i + 1]] = xi
        y[off[i]:off[i + 1]] = yi
    if nan_frac > 0:
        y[rng.random(len(y)) < nan_frac] = np.nan
    xq = np.linspace(0.45, 1.55, n_query)
    return off, x, y, xq


def make_quote_chain(seed: int, spec, device, s0: float = 4500.0, t0: int = T0_US + 34_200 * US, tick_us: int = US,
                     chunk_t: int = 128, drop_frac: float = 0.0):
    """Config 2: one SPX-like chain, ``spec.n_t`` ticks x ``spec.n_exp`` expiries x
    ``spec.n_strike`` strikes, generated on ``device`` with torch's seeded generator, time-major.

    spot: geometric random walk; strikes: a fixed ladder over [0.5, 1.5] s0; T_e(t) = (7 + 12 e) days - t;
    sigma: a skewed smile in ln(K/S)/sqrt(T) clipped to [0.05, 1.5] plus 0.3 vol-point noise; 90 % of the
    contracts are out-of-the-money at the open, 10 % their in-the-money twins (call/put is per contract); r = 0.05.  Prices come from the
    library's own forward pricer (ivx_bs_price).  ``drop_frac`` of the prices are set to NaN (missing quotes).
    Returns a surface.QuoteChain.
    """
    import torch

    from . import device as D
    from .surface import QuoteChain

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = dict(dtype=torch.float64, device=device)
    n_t, n_e, n_k = spec.n_t, spec.n_exp, spec.n_strike
    n = spec.n_quotes
    ts = t0 + torch.arange(n_t, dtype=torch.int64, device=device) * tick_us
    spot = s0 * torch.exp(torch.cumsum(torch.randn(n_t, generator=g, **f64) * 1e-4, 0))
    ladder = s0 * torch.linspace(0.5, 1.5, n_k, **f64)
    T0 = (7.0 + 12.0 * torch.arange(n_e, **f64)) / 365.0
    level = 0.2 + 0.02 * torch.sin(torch.arange(n_t, **f64) * (6.283185307179586 / max(1, n_t)))
    S = torch.empty(n, **f64); K = torch.empty(n, **f64); T = torch.empty(n, **f64)
    sigma = torch.empty(n, **f64)
    is_call = torch.empty(n, dtype=torch.uint8, device=device)
    per_t = n_e * n_k
    ke = (torch.arange(n_e, device=device)[:, None] * 3 + torch.arange(n_k, device=device)[None, :] * 7)
    # call/put is a property of the contract: out-of-the-money side at the open, every tenth contract the other side
    call_ek = ((ladder[None, :] >= s0).expand(n_e, n_k) ^ ((ke % 10) == 0)).to(torch.uint8)
    for a in range(0, n_t, chunk_t):
        b = min(n_t, a + chunk_t)
        sl = slice(a * per_t, b * per_t)
        s_ = spot[a:b, None, None].expand(b - a, n_e, n_k)
        k_ = ladder[None, None, :].expand(b - a, n_e, n_k)
        t_ = (T0[None, :, None] - (ts[a:b, None, None] - t0).to(torch.float64) / (365.0 * DAY_US)).expand(b - a, n_e, n_k)
        m = torch.log(k_ / s_) / torch.sqrt(t_)
        sg = level[a:b, None, None] + 0.35 * m * m - 0.08 * m + 0.003 * torch.randn(b - a, n_e, n_k, generator=g, **f64)
        S[sl] = s_.reshape(-1); K[sl] = k_.reshape(-1); T[sl] = t_.reshape(-1)
        sigma[sl] = sg.clamp_(0.05, 1.5).reshape(-1)
        is_call[sl] = call_ek[None].expand(b - a, n_e, n_k).reshape(-1)
        del s_, k_, t_, m, sg
    r = torch.full((n,), 0.05, **f64)
    price = D.bs_price(S, K, T, r, sigma, is_call)
    if drop_frac > 0:
        for a in range(0, n, 1 << 26):
            b = min(n, a + (1 << 26))
            price[a:b][torch.rand(b - a, generator=g, device=device) < drop_frac] = float("nan")
    grid_k = s0 * torch.linspace(0.55, 1.45, spec.n_grid_k, **f64)
    grid_m = torch.linspace(10.0 / 365.0, 340.0 / 365.0, spec.n_grid_m, **f64)
    chain = QuoteChain(spec, ts, price, S, K, T, r, is_call, grid_k, grid_m)
    chain.sigma_true = sigma
    return chain


def make_quote_chain_np(seed: int, n_t: int, n_exp: int, n_strike: int, pricer, n_grid_k: int = 512, n_grid_m: int = 64,
                        s0: float = 4500.0, t0: int = T0_US + 34_200 * US, tick_us: int = US, drop_frac: float = 0.0):
    """Host (numpy) twin of ``make_quote_chain`` — same shapes and distributions, numpy's generator — for the CPU
    baseline and the small parity cases.  ``pricer(S,K,T,r,sigma,is_call)`` supplies the prices.
    Returns a dict of arrays: ts, price, S, K, T, r, is_call, sigma_true, grid_k, grid_m."""
    rng = np.random.default_rng(seed)
    ts = t0 + np.arange(n_t, dtype=np.int64) * tick_us
    spot = s0 * np.exp(np.cumsum(rng.standard_normal(n_t) * 1e-4))
    ladder = s0 * np.linspace(0.5, 1.5, n_strike)
    T0 = (7.0 + 12.0 * np.arange(n_exp)) / 365.0
    level = 0.2 + 0.02 * np.sin(np.arange(n_t) * (2 * np.pi / max(1, n_t)))
    shape = (n_t, n_exp, n_strike)
    S = np.broadcast_to(spot[:, None, None], shape)
    K = np.broadcast_to(ladder[None, None, :], shape)
    T = np.broadcast_to(T0[None, :, None] - ((ts - t0) / (365.0 * DAY_US))[:, None, None], shape)
    m = np.log(K / S) / np.sqrt(T)
    sigma = np.clip(level[:, None, None] + 0.35 * m * m - 0.08 * m + 0.003 * rng.standard_normal(shape), 0.05, 1.5)
    ke = np.arange(n_exp)[:, None] * 3 + np.arange(n_strike)[None, :] * 7
    call_ek = (np.broadcast_to(ladder[None, :] >= s0, (n_exp, n_strike)) ^ ((ke % 10) == 0))
    is_call = np.broadcast_to(call_ek[None], shape).astype(np.uint8)
    flat = lambda a: np.ascontiguousarray(a).reshape(-1)
    S, K, T, sigma, is_call = flat(S), flat(K), flat(T), flat(sigma), flat(is_call)
    r = np.full(S.shape, 0.05)
    price = np.asarray(pricer(S, K, T, r, sigma, is_call), dtype=np.float64)
    if drop_frac > 0:
        price[rng.random(price.shape) < drop_frac] = np.nan
    return dict(ts=ts, price=price, S=S, K=K, T=T, r=r, is_call=is_call, sigma_true=sigma,
                grid_k=s0 * np.linspace(0.55, 1.45, n_grid_k), grid_m=np.linspace(10.0 / 365.0, 340.0 / 365.0, n_grid_m))
