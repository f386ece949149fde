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
