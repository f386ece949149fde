"""The north_star pipeline for one option chain (one underlying, one trading day):

    quotes --(kernel C: Black-Scholes implied vol, one thread per quote)--> iv[t][expiry][strike]
           --(kernel D: natural cubic spline across strike, per (timestamp, expiry) slice)--> smile[t][expiry][grid_k]
           --(kernel D': natural cubic spline across maturity, per (timestamp, grid strike))--> surface[t][grid_m][grid_k]
           --(kernel B': OHLC over epoch-anchored time buckets, per surface point)--> candles[bucket][grid_m][grid_k]

The reference has no such code: its conventions come from reference
src/interpolation/greeks.py:21-22 (pricing) and src/candle_reconstruction/core.py:68-89 (bucketing and the
first/max/min/last aggregation, complete-bucket rule).  All compute is in libivx.so; torch supplies memory
and streams.  Quotes are laid out time-major, ``q = (t * n_exp + e) * n_strike + k``, so each
(timestamp, expiry) smile is one contiguous slice of the quote columns and no reshuffle is needed
between the stages.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import device as D
from ._lib import IvxError

QUOTE_F64 = ("price", "S", "K", "T", "r")


@dataclass
class ChainSpec:
    n_t: int                      # timestamps (ticks) of the day
    n_exp: int                    # expiries
    n_strike: int                 # quoted strikes per expiry
    n_grid_k: int = 512           # output strike grid
    n_grid_m: int = 64            # output maturity grid
    bucket_ticks: int = 60_000_000   # candle width in ts ticks (1 min at datetime64[us])
    min_count: int = 1            # a bucket is kept iff it holds at least this many stamps (candle core.py:87-89)
    extrap: int = 2               # outside the quoted strikes/maturities: clamp flat

    @property
    def n_quotes(self) -> int:
        return self.n_t * self.n_exp * self.n_strike

    @property
    def n_series(self) -> int:
        return self.n_grid_m * self.n_grid_k


@dataclass
class QuoteChain:
    """Device-resident quotes of one chain, structure of arrays (49 B per quote with the IV output)."""
    spec: ChainSpec
    ts: torch.Tensor              # int64[n_t]
    price: torch.Tensor           # float64[n_quotes]
    S: torch.Tensor
    K: torch.Tensor
    T: torch.Tensor
    r: torch.Tensor
    is_call: torch.Tensor         # uint8[n_quotes]
    grid_k: torch.Tensor          # float64[n_grid_k] strikes of the output grid, ascending
    grid_m: torch.Tensor          # float64[n_grid_m] maturities (years) of the output grid, ascending

    def check(self) -> None:
        n = self.spec.n_quotes
        for name in QUOTE_F64:
            t = getattr(self, name)
            if t.dtype != torch.float64 or t.numel() != n:
                raise IvxError(f"QuoteChain.{name}: expected float64[{n}]")
        if self.is_call.dtype != torch.uint8 or self.is_call.numel() != n:
            raise IvxError(f"QuoteChain.is_call: expected uint8[{n}]")
        if self.ts.dtype != torch.int64 or self.ts.numel() != self.spec.n_t:
            raise IvxError("QuoteChain.ts: expected int64[n_t]")
        if self.grid_k.numel() != self.spec.n_grid_k or self.grid_m.numel() != self.spec.n_grid_m:
            raise IvxError("QuoteChain grids do not match the spec")


@dataclass
class StructuredChain:
    """The same chain with its structure spelled out instead of repeated per quote: the underlying and the rate
    are per stamp, the strike per contract, the time to expiry per (stamp, expiry).  9 B per quote instead of 41."""
    spec: ChainSpec
    ts: torch.Tensor              # int64[n_t]
    price: torch.Tensor           # float64[n_quotes], [t][e][k]
    is_call: torch.Tensor         # uint8[n_exp * n_strike] (call/put is a property of the contract) or uint8[n_quotes]
    S_t: torch.Tensor             # float64[n_t]
    K_ek: torch.Tensor            # float64[n_exp * n_strike], ascending inside an expiry
    T_te: torch.Tensor            # float64[n_t * n_exp], ascending inside a stamp
    r_t: torch.Tensor             # float64[n_t]
    grid_k: torch.Tensor
    grid_m: torch.Tensor

    @staticmethod
    def from_quotes(q: "QuoteChain") -> "StructuredChain":
        """Strip the redundancy of a per-quote chain (tests and benches; assumes the chain IS structured)."""
        sp = q.spec
        v = lambda t: t.view(sp.n_t, sp.n_exp, sp.n_strike)
        call3 = v(q.is_call)
        per_contract = bool((call3 == call3[:1]).all().item())
        is_call = call3[0].reshape(-1).contiguous() if per_contract else q.is_call
        return StructuredChain(sp, q.ts, q.price, is_call, v(q.S)[:, 0, 0].contiguous(), v(q.K)[0].reshape(-1).contiguous(),
                               v(q.T)[:, :, 0].reshape(-1).contiguous(), v(q.r)[:, 0, 0].contiguous(), q.grid_k, q.grid_m)


@dataclass
class SurfaceResult:
    iv: torch.Tensor              # float64[n_quotes]
    smile: torch.Tensor           # float64[n_t * n_exp * n_grid_k]
    surface: torch.Tensor         # float64[n_t * n_grid_m * n_grid_k]
    bucket_ts: torch.Tensor       # int64[n_buckets]
    count: torch.Tensor           # int32[n_buckets] stamps per bucket
    open: torch.Tensor            # float64[n_buckets * n_series]; NaN rows where count < min_count
    high: torch.Tensor
    low: torch.Tensor
    close: torch.Tensor
    n_buckets: int


class SurfacePipeline:
    """Owns the intermediate buffers of one chain shape so that a step allocates nothing."""

    STAGES = ("implied_vol", "spline_strike", "spline_maturity", "candles")

    def __init__(self, spec: ChainSpec, device: torch.device, max_buckets: Optional[int] = None, banded=True):
        """``banded``: True = strike spline through the banded operators (+ exact solve of the slices they cannot take); False =
        scan kernel; "auto" = banded for the first run, then the share of slices that needed the exact solve is read back ONCE
        and above 25 % (a tape with missing quotes nearly everywhere) the pipeline stays on the scan kernel — the two paths
        differ in the last bits, which is why this is opt-in."""
        if device.type != "cuda":
            raise IvxError("SurfacePipeline needs a CUDA device (no CPU path exists)")
        self.spec = spec
        self.device = device
        f64 = dict(dtype=torch.float64, device=device)
        self.iv = torch.empty(spec.n_quotes, **f64)
        self.smile = torch.empty(spec.n_t * spec.n_exp * spec.n_grid_k, **f64)
        self.surface = torch.empty(spec.n_t * spec.n_series, **f64)
        self.knot_off = torch.arange(spec.n_t * spec.n_exp + 1, dtype=torch.int64, device=device) * spec.n_strike
        self.max_buckets = max_buckets if max_buckets is not None else spec.n_t
        self.candles = [torch.empty(self.max_buckets * spec.n_series, **f64) for _ in range(4)]
        self.count = torch.empty(self.max_buckets, dtype=torch.int32, device=device)
        self._ws = torch.empty(256, dtype=torch.uint8, device=device)
        self._buckets = None       # (ts tensor identity, bucket_ts, bucket_start, n_buckets)
        # chain layout: every slice (t, e) reads its abscissae from the shared ladder of expiry e
        self._x_off = None
        # strike spline on shared ladders: banded operators (ivx_spline_band.cu) when the shapes allow, else the scan kernel
        self.banded_auto = banded == "auto"
        self.banded = bool(banded) and spec.n_grid_k <= 512 and 48 <= spec.n_strike <= 2048
        self._band = self._q_band = None
        self._ivl_ws = None
        self._order = None
        self._fb_host = torch.zeros(1, dtype=torch.int32).pin_memory() if self.banded and self.banded_auto else None
        self._fb_event = None
        self.banded_switched_off = False
        self._chain_ws = None
        self._iv_ws = None
        self._lu_key = None
        self._q_lu = None
        self._q_lu_key = None

    def _group_slices(self):
        """Slices (t, e) grouped by ladder e (CSR), once: slice id = t * n_exp + e."""
        if self._order is None:
            sp = self.spec
            e = torch.arange(sp.n_exp, dtype=torch.int64, device=self.device)
            t = torch.arange(sp.n_t, dtype=torch.int64, device=self.device)
            self._order = (t[None, :] * sp.n_exp + e[:, None]).reshape(-1).to(torch.int32).contiguous()
            self._lad_slice_off = torch.arange(sp.n_exp + 1, dtype=torch.int64, device=self.device) * sp.n_t
            self._fb = (torch.empty(sp.n_t * sp.n_exp, dtype=torch.int32, device=self.device),
                        torch.zeros(1, dtype=torch.int32, device=self.device))

    def plan_buckets(self, ts: torch.Tensor):
        """Bucket boundaries of the day's stamps (one tiny kernel + an 8-byte read-back, done once per tape)."""
        _, bucket_ts, bucket_start, nb = D.time_buckets(ts, self.spec.bucket_ticks)
        if nb > self.max_buckets:
            raise IvxError(f"{nb} buckets exceed the planned {self.max_buckets}")
        self._buckets = (ts.data_ptr(), bucket_ts, bucket_start, nb)
        return self._buckets

    def _implied_vols(self, q: "QuoteChain", structured: bool):
        """Stage 1 of ``run``: implied vols of every quote into ``self.iv`` (and, on the banded per-quote path, the slices whose
        strikes differ from their ladder's into ``self._fb``)."""
        sp = self.spec
        if structured:
            if self._chain_ws is None:
                from ._lib import lib
                self._chain_ws = torch.empty(lib().ivx_bs_chain_workspace_bytes(sp.n_t, sp.n_exp, sp.n_strike), dtype=torch.uint8,
                                             device=self.device)
                self._x_off = (torch.arange(sp.n_t * sp.n_exp, dtype=torch.int64, device=self.device) % sp.n_exp) * sp.n_strike
                self._ladder_off = torch.arange(sp.n_exp + 1, dtype=torch.int64, device=self.device) * sp.n_strike
                n_tab = lib().ivx_spline_ladder_table_len(sp.n_exp, sp.n_strike)
                self._lu = [torch.zeros(n_tab, dtype=torch.float64, device=self.device) for _ in range(3)]
                self._slice_ladder = (torch.arange(sp.n_t * sp.n_exp, dtype=torch.int64, device=self.device) % sp.n_exp).to(torch.int32)
            D.bs_implied_vol_chain(q.price, q.is_call, q.S_t, q.K_ek, q.T_te, q.r_t, n_t=sp.n_t, n_exp=sp.n_exp,
                                   n_strike=sp.n_strike, out=self.iv, workspace=self._chain_ws)
        else:
            if self.banded:
                # the solver has every strike in a register: it compares them with the first stamp's ladders on the way and flags
                # the slices that differ, so the strike spline below does not read the K column again (x_verified)
                self._group_slices()
                if self._ivl_ws is None:
                    from ._lib import lib
                    self._ivl_ws = torch.empty(lib().ivx_bs_implied_vol_ladder_workspace_bytes(sp.n_t, sp.n_exp * sp.n_strike), dtype=torch.uint8,
                                               device=self.device)
                D.bs_implied_vol_ladder(q.price, q.S, q.K, q.T, q.r, q.is_call, n_t=sp.n_t, period=sp.n_exp * sp.n_strike, n_strike=sp.n_strike,
                                        x_ladder=q.K, flags=self._fb, out=self.iv, workspace=self._ivl_ws)
            else:
                if self._iv_ws is None:
                    from ._lib import lib
                    self._iv_ws = torch.empty(lib().ivx_bs_implied_vol_workspace_bytes(sp.n_quotes), dtype=torch.uint8, device=self.device)
                D.bs_implied_vol(q.price, q.S, q.K, q.T, q.r, q.is_call, out=self.iv, workspace=self._iv_ws)

    def run(self, q: QuoteChain, events: Optional[Dict[str, List[torch.cuda.Event]]] = None, part: str = "all") -> SurfaceResult:
        """One pass of the hot path.  ``events`` (stage -> [start, end]) are recorded on the current stream.
        ``part``: "all", or "iv" (stage 1 only, returns None) / "rest" (stages 2-4 on the IVs of an earlier "iv" call) — the two
        halves of a chunked day can then run on two streams, the solver of one chunk beside the splines of the chunk before."""
        sp = self.spec
        if self._buckets is None or self._buckets[0] != q.ts.data_ptr():
            self.plan_buckets(q.ts)
        _, bucket_ts, bucket_start, nb = self._buckets

        def mark(stage, k):
            if events is not None:
                events[stage][k].record()

        structured = isinstance(q, StructuredChain)
        if self._fb_event is not None:                   # "auto": decided once, after the first run
            self._fb_event.synchronize()
            self._fb_event, self.banded_auto = None, False
            if int(self._fb_host.item()) > 0.25 * sp.n_t * sp.n_exp:
                self.banded, self.banded_switched_off = False, True
        if part != "rest":
            mark("implied_vol", 0)
            self._implied_vols(q, structured)
            mark("implied_vol", 1)
        if part == "iv":
            return None
        mark("spline_strike", 0)
        # knots of slice (t, e): x = the quoted strikes, read in place from the K column / the expiry's ladder
        if structured:
            # the x-only half of the elimination once per expiry ladder — and only when the ladders changed (contracts do
            # not change intraday) — then right-hand sides only
            key = (q.K_ek.data_ptr(), q.K_ek._version, q.grid_k.data_ptr(), q.grid_k._version)
            if self._lu_key != key:
                D.spline_ladder_factor(self._ladder_off, q.K_ek, max_knots=sp.n_strike, out=self._lu)
                self._band = D.spline_band_build(self._ladder_off, q.K_ek, q.grid_k, extrap=sp.extrap, out=self._band,
                                                 n_ladder_knots=sp.n_exp * sp.n_strike) if self.banded else None
                self._lu_key = key
            if self.banded:
                # every slice is n_grid_k dot products of 48 terms against its ladder's banded operator; slices with a missing
                # knot take the exact scan right behind (same call)
                self._group_slices()
                D.spline_fit_eval_banded(self.knot_off, self._x_off, self._slice_ladder, self._order, self._lad_slice_off, sp.n_t,
                                         self._ladder_off, q.K_ek, self._lu, self.iv, q.grid_k, self._band, max_knots=sp.n_strike,
                                         out=self.smile, fb=self._fb)
            else:
                D.spline_fit_eval(self.knot_off, q.K_ek, self.iv, q.grid_k, max_knots=sp.n_strike, extrap=sp.extrap,
                                  out=self.smile, x_off=self._x_off, lu=self._lu, slice_ladder=self._slice_ladder)
        else:
            # per-quote strikes: a chain's strike ladders normally do not change during the day, but nothing says so here.
            # The first stamp's ladders are factored (30 small launches' worth of work) and every slice is PRESUMED to
            # sit on its expiry's; the kernel verifies that bit for bit while staging and solves the others in full.
            if self._q_lu is None:
                from ._lib import lib
                n_tab = lib().ivx_spline_ladder_table_len(sp.n_exp, sp.n_strike)
                self._q_lu = [torch.zeros(n_tab, dtype=torch.float64, device=self.device) for _ in range(3)]
                self._q_x_off = (torch.arange(sp.n_t * sp.n_exp, dtype=torch.int64, device=self.device) % sp.n_exp) * sp.n_strike
                self._q_ladder_off = torch.arange(sp.n_exp + 1, dtype=torch.int64, device=self.device) * sp.n_strike
                self._q_slice_ladder = (torch.arange(sp.n_t * sp.n_exp, dtype=torch.int64, device=self.device) % sp.n_exp).to(torch.int32)
            key = (q.K.data_ptr(), q.K._version, q.grid_k.data_ptr(), q.grid_k._version)   # tables follow the columns: rebuilt whenever one is written
            if self._q_lu_key != key:
                D.spline_ladder_factor(self._q_ladder_off, q.K, max_knots=sp.n_strike, out=self._q_lu)
                self._q_band = D.spline_band_build(self._q_ladder_off, q.K, q.grid_k, extrap=sp.extrap, out=self._q_band,
                                                   n_ladder_knots=sp.n_exp * sp.n_strike) if self.banded else None
                self._q_lu_key = key
            if self.banded:
                # the first stamp's ladders as banded operators; a slice is taken that way only if its own strikes equal its
                # ladder's bit for bit (checked while its knots are staged), the exact scan solves the others right behind
                self._group_slices()
                D.spline_fit_eval_banded(self.knot_off, self._q_x_off, self._q_slice_ladder, self._order, self._lad_slice_off, sp.n_t,
                                         self._q_ladder_off, q.K, self._q_lu, self.iv, q.grid_k, self._q_band, max_knots=sp.n_strike,
                                         x_own=q.K, out=self.smile, fb=self._fb, x_verified=True)
            else:
                D.spline_fit_eval(self.knot_off, q.K, self.iv, q.grid_k, max_knots=sp.n_strike, extrap=sp.extrap, out=self.smile,
                                  x_off=self._q_x_off, lu=self._q_lu, slice_ladder=self._q_slice_ladder, x_own=q.K,
                                  ladder_off=self._q_ladder_off)
        if self.banded and self.banded_auto:
            self._fb_host.copy_(self._fb[1], non_blocking=True)
            self._fb_event = torch.cuda.Event()
            self._fb_event.record()
        mark("spline_strike", 1)
        mark("spline_maturity", 0)
        # plane t: n_grid_k systems with n_exp knots at x = T[t][e], read in place
        if structured:
            D.spline_strided_batched(q.T_te, sp.n_exp, 1, self.smile, q.grid_m, sp.n_t, sp.n_grid_k, sp.n_exp,
                                     extrap=sp.extrap, out=self.surface)
        else:
            D.spline_strided_batched(q.T, sp.n_exp * sp.n_strike, sp.n_strike, self.smile, q.grid_m,
                                     sp.n_t, sp.n_grid_k, sp.n_exp, extrap=sp.extrap, out=self.surface)
        mark("spline_maturity", 1)
        mark("candles", 0)
        D.ohlc_time_major(self.surface, sp.n_series, bucket_start, nb, sp.min_count, outs=self.candles, count=self.count)
        mark("candles", 1)
        m = nb * sp.n_series
        return SurfaceResult(self.iv, self.smile, self.surface, bucket_ts, self.count[:nb],
                             self.candles[0][:m], self.candles[1][:m], self.candles[2][:m], self.candles[3][:m], nb)


# ------------------------------------------------------------------------------------------------------
# Host-facing entry: quotes in (pinned) host memory -> candles in host memory
# ------------------------------------------------------------------------------------------------------
@dataclass
class HostQuoteChunk:
    """A run of whole time buckets of one chain in host memory (pinned for full PCIe speed), per-quote layout:
    ``ts`` int64[n_t], the five float64 quote columns and ``is_call`` uint8, each n_t*n_exp*n_strike long."""
    ts: torch.Tensor
    price: torch.Tensor
    S: torch.Tensor
    K: torch.Tensor
    T: torch.Tensor
    r: torch.Tensor
    is_call: torch.Tensor
    FIELDS = ("ts", "price", "S", "K", "T", "r", "is_call")

    @property
    def n_t(self) -> int:
        return self.ts.numel()

    def nbytes(self) -> int:
        return sum(getattr(self, f).numel() * getattr(self, f).element_size() for f in self.FIELDS)


@dataclass
class HostChainChunk:
    """The same run in the chain layout: per-quote ``price``; ``S_t``, ``r_t`` float64[n_t]; ``T_te`` float64[n_t*n_exp].
    The strike ladder ``K_ek`` and the call/put flags ``is_call_ek`` belong to the stream (contracts do not change
    intraday): 8 B per quote cross PCIe."""
    ts: torch.Tensor
    price: torch.Tensor
    S_t: torch.Tensor
    T_te: torch.Tensor
    r_t: torch.Tensor
    FIELDS = ("ts", "price", "S_t", "T_te", "r_t")

    @property
    def n_t(self) -> int:
        return self.ts.numel()

    def nbytes(self) -> int:
        return sum(getattr(self, f).numel() * getattr(self, f).element_size() for f in self.FIELDS)


@dataclass
class HostTickChainChunk(HostChainChunk):
    """The chain layout with ``price`` as int32 ticks of 1e-4 — the precision the reference stores prices at (DECIMAL(15,4),
    reference src/database/schema.py:30-45) — ``INT32_MIN`` = no quote: 4 B per quote cross PCIe.  An explicit, separately named
    wire format: the IVs are those of the prices ROUNDED to four decimals (bit-equal to solving ``round(price, 4)``)."""

    @staticmethod
    def encode(price: torch.Tensor) -> torch.Tensor:
        """float64 prices -> int32 ticks (round half to even at 1e-4; NaN -> INT32_MIN; prices must stay below 214 748.3647)."""
        t = torch.round(price * 1e4)
        if bool((t.abs() >= 2 ** 31).any()):
            raise IvxError("HostTickChainChunk.encode: a price does not fit int32 ticks of 1e-4")
        return torch.where(torch.isnan(price), torch.full_like(t, -2.0 ** 31), t).to(torch.int32)


class HostSurfaceStream:
    """Streams a day of quotes from host memory through the pipeline, chunk by chunk, double-buffered:
    while chunk i is being solved on the compute stream, chunk i+1 is on its way over PCIe on the copy
    stream, and the candles of chunk i-1 are on their way back.  A chunk must hold whole time buckets
    (the caller cuts the tape at bucket boundaries) and exactly ``chunk_spec.n_t`` stamps.  ``K_ek`` selects
    the chain layout (HostChainChunk, with ``is_call_ek``); without it chunks are per-quote (HostQuoteChunk).
    """

    def __init__(self, chunk_spec: ChainSpec, device: torch.device, grid_k: torch.Tensor, grid_m: torch.Tensor,
                 max_buckets_per_chunk: int, K_ek: Optional[torch.Tensor] = None, is_call_ek: Optional[torch.Tensor] = None,
                 price_ticks: bool = False):
        self.spec = chunk_spec
        self.device = device
        self.grid_k = grid_k.to(device)
        self.grid_m = grid_m.to(device)
        self.K_ek = None if K_ek is None else K_ek.to(device)
        if (K_ek is None) != (is_call_ek is None):
            raise IvxError("the chain layout needs both K_ek and is_call_ek")
        self.is_call_ek = None if is_call_ek is None else is_call_ek.to(device)
        if price_ticks and K_ek is None:
            raise IvxError("the int32 tick price format belongs to the chain layout (pass K_ek and is_call_ek)")
        self.chunk_type = HostQuoteChunk if K_ek is None else (HostTickChainChunk if price_ticks else HostChainChunk)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.out_stream = torch.cuda.Stream(device=device)
        self.max_buckets = max_buckets_per_chunk
        sp = chunk_spec
        sizes = {"ts": (sp.n_t, torch.int64), "price": (sp.n_quotes, torch.int32 if price_ticks else torch.float64), "is_call": (sp.n_quotes, torch.uint8),
                 "S": (sp.n_quotes, torch.float64), "K": (sp.n_quotes, torch.float64), "T": (sp.n_quotes, torch.float64),
                 "r": (sp.n_quotes, torch.float64), "S_t": (sp.n_t, torch.float64), "r_t": (sp.n_t, torch.float64),
                 "T_te": (sp.n_t * sp.n_exp, torch.float64)}
        self.slots = []
        for _ in range(2):
            cols = {f: torch.empty(sizes[f][0], dtype=sizes[f][1], device=device) for f in self.chunk_type.FIELDS}
            self.slots.append(dict(cols=cols, pipe=SurfacePipeline(chunk_spec, device, max_buckets=max_buckets_per_chunk),
                                   loaded=torch.cuda.Event(), computed=torch.cuda.Event(), drained=torch.cuda.Event()))
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run(self, chunks: List, out_ohlc: List[torch.Tensor], out_bucket_ts: torch.Tensor, out_count: torch.Tensor,
            bucket_counts: Optional[List[int]] = None) -> int:
        """Process ``chunks`` in order; candles land in the (pinned) host tensors ``out_ohlc`` (4 x
        float64[total_buckets * n_series]), bucket stamps/counts in ``out_bucket_ts``/``out_count``.
        ``bucket_counts[i]`` = number of time buckets in chunk i (computed from the host stamps when omitted).
        Returns the number of buckets written.  Synchronises once, at the end."""
        sp = self.spec
        compute = torch.cuda.current_stream(self.device)
        n_series = sp.n_series
        row = 0
        self.h2d_bytes = self.d2h_bytes = 0
        for slot in self.slots:
            slot["drained"].record(compute)
        for i, ch in enumerate(chunks):
            if type(ch) is not self.chunk_type:
                raise IvxError(f"chunk {i} is a {type(ch).__name__}, this stream takes {self.chunk_type.__name__}")
            if ch.n_t != sp.n_t:
                raise IvxError(f"chunk {i} has {ch.n_t} stamps, the stream was planned for {sp.n_t} "
                               "(build a second stream for the ragged tail)")
            slot = self.slots[i & 1]
            # the bucket layout of a chunk is host knowledge: the stamps are host data
            nb = bucket_counts[i] if bucket_counts is not None else int(torch.unique_consecutive(ch.ts // sp.bucket_ticks).numel())
            if nb > self.max_buckets:
                raise IvxError(f"chunk {i} spans {nb} buckets (> {self.max_buckets})")
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(slot["drained"])          # the slot's previous results have left
                for name in ch.FIELDS:
                    slot["cols"][name].copy_(getattr(ch, name), non_blocking=True)
                slot["loaded"].record(self.copy_stream)
            self.h2d_bytes += ch.nbytes()
            compute.wait_event(slot["loaded"])
            c = slot["cols"]
            if self.K_ek is None:
                q = QuoteChain(sp, c["ts"], c["price"], c["S"], c["K"], c["T"], c["r"], c["is_call"], self.grid_k, self.grid_m)
            else:
                q = StructuredChain(sp, c["ts"], c["price"], self.is_call_ek, c["S_t"], self.K_ek, c["T_te"], c["r_t"],
                                    self.grid_k, self.grid_m)
            pipe = slot["pipe"]
            # time_buckets reads its count back (8 bytes): the host already knows nb, so plan without the sync
            _, bucket_ts, _, _ = _plan_buckets_nosync(pipe, c["ts"], nb)
            res = pipe.run(q)
            slot["computed"].record(compute)
            with torch.cuda.stream(self.out_stream):
                self.out_stream.wait_event(slot["computed"])
                m = nb * n_series
                for k, src in enumerate((res.open, res.high, res.low, res.close)):
                    out_ohlc[k][row * n_series: row * n_series + m].copy_(src[:m], non_blocking=True)
                out_bucket_ts[row: row + nb].copy_(bucket_ts[:nb], non_blocking=True)
                out_count[row: row + nb].copy_(pipe.count[:nb], non_blocking=True)
                slot["drained"].record(self.out_stream)
            self.d2h_bytes += 4 * m * 8 + nb * 12
            row += nb
        compute.wait_stream(self.out_stream)
        compute.synchronize()
        return row


def _plan_buckets_nosync(pipe: SurfacePipeline, ts: torch.Tensor, nb: int):
    n_t = ts.numel()
    if not hasattr(pipe, "_bk"):
        dev = ts.device
        pipe._bk = (torch.empty(n_t, dtype=torch.int32, device=dev), torch.empty(n_t, dtype=torch.int64, device=dev),
                    torch.empty(n_t + 1, dtype=torch.int32, device=dev), torch.empty(1, dtype=torch.int64, device=dev))
    bucket_of_t, bucket_ts, bucket_start, nbuf = pipe._bk
    D.time_buckets_into(ts, pipe.spec.bucket_ticks, bucket_of_t, bucket_ts, bucket_start, nbuf)
    pipe._buckets = (ts.data_ptr(), bucket_ts, bucket_start, nb)
    return pipe._buckets
