"""Device-level operators: torch CUDA tensors in, torch CUDA tensors out, all compute in libivx.so.

PyTorch is plumbing here (device memory, streams); every kernel is a hand-written sm_100a kernel
reached through the C ABI of include/ivx.h.  Nothing in this module computes on the CPU, and
nothing falls back: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import check, lib, ptr_array

_LAUNCHES = 0     # kernels of ours enqueued since the last reset (bench.py reports it)

# kernels each entry point enqueues (memsets are not kernels of ours and are not counted); ohlcv_reduce: tile plan, reduce, finalize
_KERNELS_PER_CALL = {
    "ohlcv_reduce": 3, "time_buckets": 1, "ohlc_time_major": 1, "interp_plan": 3, "interp_exec": 1,
    "bs_implied_vol": 1, "bs_implied_vol_2k": 2, "bs_implied_vol_ladder": 2, "bs_price": 1, "bs_greeks": 1, "spline_fit_eval": 1, "spline_strided": 1,
    "spline_strided_batched": 1, "scan": 1, "probe": 1, "bs_implied_vol_chain": 3, "spline_ladder_factor": 1,
    "spline_band_build": 6, "spline_fit_eval_banded": 2, "spline_fit_eval_banded_mma": 3, "bridge_quality": 1, "gather": 1, "random_tape": 1, "bridge_plan": 4, "bridge_ohlcv": 1, "bridge_ohlcv_seg": 3, "read_back": 1,   # plan: count, 2 scans, offsets; segments, check, (idle) fallback
}


def reset_launch_count() -> None:
    global _LAUNCHES
    _LAUNCHES = 0


def launch_count() -> int:
    return _LAUNCHES


def _count(name: str) -> None:
    global _LAUNCHES
    _LAUNCHES += _KERNELS_PER_CALL[name]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.IvxError("libivx operators take CUDA tensors (no CPU path exists)")
    if not t.is_contiguous():
        raise _lib.IvxError("libivx operators take contiguous tensors")
    return t.data_ptr()


def read_back(vals: torch.Tensor) -> List[int]:
    """A few int64 sizes from the device to the host without the copy engine (``ivx_read_back_i64``): an 8-byte ``.item()`` is a
    cudaMemcpy and queues behind every bulk device-to-host copy in flight on other streams — in a streaming pipeline the plan
    of batch k+1 then waits for the whole drain of batch k.  Blocks until the current stream has produced ``vals``."""
    import ctypes
    n = vals.numel()
    if vals.dtype != torch.int64 or not vals.is_contiguous():
        vals = vals.to(torch.int64).contiguous()
    out = (ctypes.c_int64 * max(1, n))()
    check(lib().ivx_read_back_i64(n, _ptr(vals), out, _stream()), "ivx_read_back_i64")
    _count("read_back")
    return [int(out[i]) for i in range(n)]


def _req(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if t.dtype != dtype:
        raise _lib.IvxError(f"{name}: expected {dtype}, got {t.dtype}")
    return t


def _empty(n: int, dtype: torch.dtype, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(int(n), dtype=dtype, device=like.device)


# ------------------------------------------------------------------------------------------------
# Kernel B
# ------------------------------------------------------------------------------------------------
@dataclass
class CandleBatch:
    out_off: torch.Tensor          # int64[n_sym+1]
    ts: torch.Tensor               # int64[m] bucket start
    open: torch.Tensor
    high: torch.Tensor
    low: torch.Tensor
    close: torch.Tensor
    volume: torch.Tensor
    count: torch.Tensor            # int32[m] raw rows per candle
    sym_status: torch.Tensor       # uint8[n_sym]
    sym_flags: torch.Tensor        # int32[n_sym] (bit flags)
    sym_row: Optional[torch.Tensor] = None     # int64[m]
    group: Optional[torch.Tensor] = None       # int64[m]
    sym_group_off: Optional[torch.Tensor] = None
    total: int = 0


def ohlcv_reduce(in_off: torch.Tensor, ts: torch.Tensor, o: torch.Tensor, h: torch.Tensor, l: torch.Tensor,
                 c: torch.Tensor, v: torch.Tensor, *, bucket_ticks: int, min_count: int, check_len: bool = True,
                 sym_valid: Optional[torch.Tensor] = None, want_index: bool = False,
                 capacity: Optional[int] = None, trim: bool = True) -> CandleBatch:
    """CSR batch of symbols -> compacted candles (reference candle_reconstruction/core.py:59-89).

    ``trim`` reads the candle count back (one 8-byte D2H sync) and slices the outputs to it.
    """
    n_sym = in_off.numel() - 1
    n_rows = ts.numel()
    _req(in_off, torch.int64, "in_off"); _req(ts, torch.int64, "ts")
    for name, t in (("open", o), ("high", h), ("low", l), ("close", c), ("volume", v)):
        _req(t, torch.float64, name)
        if t.numel() != n_rows:
            raise _lib.IvxError(f"{name}: length {t.numel()} != {n_rows}")
    if min_count < 0 or bucket_ticks < 1:
        raise _lib.IvxError("bucket_ticks must be >= 1 and min_count >= 0")
    if capacity is None:
        capacity = n_rows // max(1, min_count) if min_count >= 1 else n_rows
    L = lib()
    ws_bytes = L.ivx_ohlcv_workspace_bytes(n_sym, n_rows)
    ws = _empty(ws_bytes, torch.uint8, ts)
    out_off = _empty(n_sym + 1, torch.int64, ts)
    o_ts = _empty(capacity, torch.int64, ts)
    oo, oh, ol, oc, ov = (_empty(capacity, torch.float64, ts) for _ in range(5))
    o_count = _empty(capacity, torch.int32, ts)
    o_sym_row = _empty(capacity, torch.int64, ts) if want_index else None
    o_group = _empty(capacity, torch.int64, ts) if want_index else None
    sym_group_off = _empty(n_sym + 1, torch.int64, ts) if want_index else None
    sym_status = _empty(n_sym, torch.uint8, ts)
    sym_flags = _empty(n_sym, torch.int32, ts)
    total = _empty(1, torch.int64, ts)
    check(L.ivx_ohlcv_reduce(
        n_sym, _ptr(in_off), n_rows, _ptr(ts), _ptr(o), _ptr(h), _ptr(l), _ptr(c), _ptr(v), _ptr(sym_valid),
        int(bucket_ticks), int(min_count), int(bool(check_len)), int(capacity), _ptr(out_off),
        _ptr(o_ts), _ptr(oo), _ptr(oh), _ptr(ol), _ptr(oc), _ptr(ov), _ptr(o_count), _ptr(o_sym_row), _ptr(o_group),
        _ptr(sym_group_off), _ptr(sym_status), _ptr(sym_flags), _ptr(total), _ptr(ws), ws_bytes, _stream()),
        "ivx_ohlcv_reduce")
    _count("ohlcv_reduce")
    res = CandleBatch(out_off, o_ts, oo, oh, ol, oc, ov, o_count, sym_status, sym_flags, o_sym_row, o_group,
                      sym_group_off)
    res._total_dev = total
    if trim:
        m = read_back(total)[0]
        if m > capacity:
            raise _lib.IvxError(f"ohlcv_reduce: {m} candles exceed the capacity {capacity}")
        res.total = m
        for name in ("ts", "open", "high", "low", "close", "volume", "count", "sym_row", "group"):
            t = getattr(res, name)
            if t is not None:
                setattr(res, name, t[:m])
    return res


def time_buckets(ts: torch.Tensor, bucket_ticks: int):
    """Stamps of one step -> (bucket_of_t int32[n_t], bucket_ts int64[nb], bucket_start int32[nb+1], nb)."""
    n_t = ts.numel()
    bucket_of_t = _empty(n_t, torch.int32, ts)
    bucket_ts = _empty(n_t, torch.int64, ts)
    bucket_start = _empty(n_t + 1, torch.int32, ts)
    nb = _empty(1, torch.int64, ts)
    check(lib().ivx_time_buckets(n_t, _ptr(_req(ts, torch.int64, "ts")), int(bucket_ticks), _ptr(bucket_of_t),
                                 _ptr(bucket_ts), _ptr(bucket_start), _ptr(nb), _stream()), "ivx_time_buckets")
    _count("time_buckets")
    n = int(nb.item())
    return bucket_of_t, bucket_ts[:n], bucket_start[:n + 1], n


def time_buckets_into(ts: torch.Tensor, bucket_ticks: int, bucket_of_t: torch.Tensor, bucket_ts: torch.Tensor,
                      bucket_start: torch.Tensor, n_buckets: torch.Tensor) -> None:
    """``time_buckets`` into caller-owned buffers, without the read-back of the bucket count."""
    n_t = ts.numel()
    if bucket_of_t.numel() < n_t or bucket_ts.numel() < n_t or bucket_start.numel() < n_t + 1 or n_buckets.numel() < 1:
        raise _lib.IvxError("time_buckets_into: output buffers too small")
    _req(bucket_of_t, torch.int32, "bucket_of_t"); _req(bucket_ts, torch.int64, "bucket_ts")
    _req(bucket_start, torch.int32, "bucket_start"); _req(n_buckets, torch.int64, "n_buckets")
    check(lib().ivx_time_buckets(n_t, _ptr(_req(ts, torch.int64, "ts")), int(bucket_ticks), _ptr(bucket_of_t),
                                 _ptr(bucket_ts), _ptr(bucket_start), _ptr(n_buckets), _stream()), "ivx_time_buckets")
    _count("time_buckets")


def ohlc_time_major(x: torch.Tensor, n_series: int, bucket_start: torch.Tensor, n_buckets: int, min_count: int,
                    outs: Optional[Sequence[torch.Tensor]] = None, count: Optional[torch.Tensor] = None):
    """x[t*n_series+s] -> four [n_buckets*n_series] candle arrays + int32 counts per bucket."""
    n_t = x.numel() // max(1, n_series)
    if outs is None:
        outs = [_empty(n_buckets * n_series, torch.float64, x) for _ in range(4)]
    elif len(outs) != 4 or any(o.numel() < n_buckets * n_series or o.dtype != torch.float64 for o in outs):
        raise _lib.IvxError("outs must be four float64 tensors of at least n_buckets*n_series elements")
    cnt = count if count is not None else _empty(n_buckets, torch.int32, x)
    if cnt.dtype != torch.int32 or cnt.numel() < n_buckets:
        raise _lib.IvxError("count must be int32[n_buckets]")
    check(lib().ivx_ohlc_time_major(n_t, n_series, _ptr(_req(x, torch.float64, "x")), n_buckets,
                                    _ptr(_req(bucket_start, torch.int32, "bucket_start")), int(min_count),
                                    _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(outs[3]), _ptr(cnt), _stream()),
          "ivx_ohlc_time_major")
    _count("ohlc_time_major")
    return outs[0], outs[1], outs[2], outs[3], cnt


# ------------------------------------------------------------------------------------------------
# Kernel A
# ------------------------------------------------------------------------------------------------
@dataclass
class InterpBatch:
    out_off: torch.Tensor                   # int64[n_sym+1]
    sym_status: torch.Tensor                # uint8[n_sym]
    sym_flags: torch.Tensor                 # int32[n_sym]
    ts: torch.Tensor                        # int64[m]
    interp: List[torch.Tensor]              # 3 x float64[m]
    fill: List[torch.Tensor]                # n_fill x float64[m]
    idx: List[torch.Tensor]                 # n_idx x int32[m]
    src_row: Optional[torch.Tensor]
    pos: Optional[torch.Tensor]
    keep: Optional[torch.Tensor]
    is_interpolated: Optional[torch.Tensor]
    total: int = 0


@dataclass
class InterpPlan:
    """Result of ivx_interp_plan: guards, grid and interval tables (in ``ws``), output sizes."""
    src_off: torch.Tensor
    n_src: int
    n_fill: int
    n_idx: int
    step: int
    ws: torch.Tensor
    ws_bytes: int
    out_off: torch.Tensor                   # int64[n_sym+1]
    sym_status: torch.Tensor
    sym_flags: torch.Tensor
    _total: Optional[int] = None

    @property
    def total(self) -> int:
        """Rows the batch will emit (reads 8 bytes back from the device the first time)."""
        if self._total is None:
            self._total = read_back(self.out_off[-1:])[0]
        return self._total


def interp_plan(src_off: torch.Tensor, ts: torch.Tensor, interp_cols: Sequence[torch.Tensor],
                fill_cols: Sequence[torch.Tensor] = (), idx_valid: Sequence[Optional[torch.Tensor]] = (),
                *, step: int, max_span: int, max_points: int = 100_000, min_points: int = 10) -> InterpPlan:
    """Step 1 of kernel A (reference interpolation/core.py:26-55,74): reads only the source rows."""
    n_sym = src_off.numel() - 1
    n_src = ts.numel()
    _req(src_off, torch.int64, "src_off"); _req(ts, torch.int64, "ts")
    if len(interp_cols) != 3:
        raise _lib.IvxError("interp_cols must be (iv, underlying_price, time_to_maturity)")
    n_fill, n_idx = len(fill_cols), len(idx_valid)
    if n_fill > _lib.MAX_FILL_COLS or n_idx > _lib.MAX_IDX_COLS:
        raise _lib.IvxError(f"at most {_lib.MAX_FILL_COLS} fill and {_lib.MAX_IDX_COLS} index columns per call")
    for t in list(interp_cols) + list(fill_cols):
        _req(t, torch.float64, "column")
        if t.numel() != n_src:
            raise _lib.IvxError("column length mismatch")
    L = lib()
    ws_bytes = L.ivx_interp_workspace_bytes(n_sym, n_src, n_fill, n_idx)
    ws = _empty(ws_bytes, torch.uint8, ts)
    out_count = _empty(n_sym, torch.int64, ts)
    out_off = _empty(n_sym + 1, torch.int64, ts)
    sym_status = _empty(n_sym, torch.uint8, ts)
    sym_flags = _empty(n_sym, torch.int32, ts)
    check(L.ivx_interp_plan(
        n_sym, _ptr(src_off), n_src, _ptr(ts), ptr_array([_ptr(t) for t in interp_cols]),
        n_fill, ptr_array([_ptr(t) for t in fill_cols]), n_idx,
        ptr_array([None if t is None else _ptr(_req(t, torch.uint8, "idx_valid")) for t in idx_valid]),
        int(step), int(max_span), int(max_points), int(min_points), _ptr(out_count), _ptr(out_off),
        _ptr(sym_status), _ptr(sym_flags), _ptr(ws), ws_bytes, _stream()), "ivx_interp_plan")
    _count("interp_plan")
    return InterpPlan(src_off, n_src, n_fill, n_idx, int(step), ws, ws_bytes, out_off, sym_status, sym_flags)


def interp_exec(plan: InterpPlan, *, want_src_row: bool = False, want_pos: bool = False, want_keep: bool = False,
                want_is_interpolated: bool = True, want_fill: bool = True, want_idx: bool = True,
                out: Optional["InterpBatch"] = None) -> "InterpBatch":
    """Step 2 of kernel A: writes the output rows.  ``out`` reuses the buffers of an earlier result of the same
    shape (no allocation inside a timed loop)."""
    n_sym = plan.src_off.numel() - 1
    m = plan.total
    like = plan.out_off
    if out is not None and out.ts.numel() == m:
        o_ts, o_interp, o_fill, o_idx = out.ts, out.interp, out.fill, out.idx
        o_src, o_pos, o_keep, o_isint = out.src_row, out.pos, out.keep, out.is_interpolated
    else:
        o_ts = _empty(m, torch.int64, like)
        o_interp = [_empty(m, torch.float64, like) for _ in range(3)]
        o_fill = [_empty(m, torch.float64, like) for _ in range(plan.n_fill)] if want_fill else []
        o_idx = [_empty(m, torch.int32, like) for _ in range(plan.n_idx)] if want_idx else []
        o_src = _empty(m, torch.int32, like) if want_src_row else None
        o_pos = _empty(m, torch.int32, like) if want_pos else None
        o_keep = _empty(m, torch.uint8, like) if want_keep else None
        o_isint = _empty(m, torch.uint8, like) if want_is_interpolated else None
    check(lib().ivx_interp_exec(
        n_sym, _ptr(plan.src_off), plan.n_src, _ptr(plan.out_off), m, plan.step, plan.n_fill, plan.n_idx, _ptr(o_ts),
        ptr_array([_ptr(t) for t in o_interp]),
        ptr_array([_ptr(t) for t in o_fill]) if o_fill else None,
        ptr_array([_ptr(t) for t in o_idx]) if o_idx else None,
        _ptr(o_src), _ptr(o_pos), _ptr(o_keep), _ptr(o_isint),
        _ptr(plan.ws), plan.ws_bytes, _stream()), "ivx_interp_exec")
    _count("interp_exec")
    return InterpBatch(plan.out_off, plan.sym_status, plan.sym_flags, o_ts, o_interp, o_fill, o_idx, o_src, o_pos, o_keep,
                       o_isint, m)


def interp_linear_ffill(src_off: torch.Tensor, ts: torch.Tensor, interp_cols: Sequence[torch.Tensor],
                        fill_cols: Sequence[torch.Tensor] = (), idx_valid: Sequence[Optional[torch.Tensor]] = (),
                        *, step: int, max_span: int, max_points: int = 100_000, min_points: int = 10,
                        want_src_row: bool = False, want_pos: bool = False, want_keep: bool = False,
                        want_is_interpolated: bool = True) -> InterpBatch:
    """CSR batch of contracts -> minute rows (reference interpolation/core.py:26-81): plan + exec.

    ``idx_valid[0]`` is the ``symbol`` column's non-null mask (uint8 tensor, or None = all valid).
    One 8-byte D2H sync (the total row count) sits between plan and exec.
    """
    plan = interp_plan(src_off, ts, interp_cols, fill_cols, idx_valid, step=step, max_span=max_span,
                       max_points=max_points, min_points=min_points)
    return interp_exec(plan, want_src_row=want_src_row, want_pos=want_pos, want_keep=want_keep,
                       want_is_interpolated=want_is_interpolated)


# ------------------------------------------------------------------------------------------------
# Kernel C
# ------------------------------------------------------------------------------------------------
def bs_implied_vol(price, S, K, T, r, is_call, *, want_iters: bool = False, out: Optional[torch.Tensor] = None,
                   workspace: Optional[torch.Tensor] = None, two_kernels: bool = True):
    """``two_kernels``: fast kernel + retry kernel over the warps it recorded (needs a small workspace, allocated here unless
    passed); False: one launch with the fallback in line.  Same results."""
    n = price.numel()
    for t in (price, S, K, T, r):
        _req(t, torch.float64, "quote column")
        if t.numel() != n:
            raise _lib.IvxError("quote column length mismatch")
    _req(is_call, torch.uint8, "is_call")
    sigma = out if out is not None else _empty(n, torch.float64, price)
    iters = _empty(n, torch.int32, price) if want_iters else None
    ws, wsb = None, 0
    if two_kernels and n:
        wsb = lib().ivx_bs_implied_vol_workspace_bytes(n)
        ws = workspace if workspace is not None and workspace.numel() >= wsb else _empty(wsb, torch.uint8, price)
    check(lib().ivx_bs_implied_vol(n, _ptr(price), _ptr(S), _ptr(K), _ptr(T), _ptr(r), _ptr(is_call), _ptr(sigma),
                                   _ptr(iters), _ptr(ws), wsb, _stream()), "ivx_bs_implied_vol")
    _count("bs_implied_vol_2k" if ws is not None else "bs_implied_vol")
    return (sigma, iters) if want_iters else sigma


def bs_implied_vol_ladder(price, S, K, T, r, is_call, *, n_t: int, period: int, n_strike: int, x_ladder: torch.Tensor,
                          flags: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None,
                          workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``bs_implied_vol`` for quotes that form slices of ``n_strike`` knots on presumed ladders (quote t * period + j presumed to have
    the strike ``x_ladder[j]``): the strikes are compared with the ladder's on the way and ``flags`` = (int32[n_t * period / n_strike],
    int32[1]) receive the slices that differ and their number."""
    n = price.numel()
    if n != n_t * period:
        raise _lib.IvxError("bs_implied_vol_ladder: price must hold n_t * period quotes")
    for t_, nm in ((price, "price"), (S, "S"), (K, "K"), (T, "T"), (r, "r"), (x_ladder, "x_ladder")):
        _req(t_, torch.float64, nm)
    _req(is_call, torch.uint8, "is_call")
    _req(flags[0], torch.int32, "flags"); _req(flags[1], torch.int32, "flag count")
    if x_ladder.numel() < period or flags[0].numel() < n_t * (period // n_strike):
        raise _lib.IvxError("bs_implied_vol_ladder: ladder or flag array too small")
    sigma = out if out is not None else _empty(n, torch.float64, price)
    wsb = lib().ivx_bs_implied_vol_ladder_workspace_bytes(n_t, period)
    ws = workspace if workspace is not None and workspace.numel() >= wsb else torch.empty(max(1, wsb), dtype=torch.uint8, device=price.device)
    check(lib().ivx_bs_implied_vol_ladder(n_t, period, n_strike, _ptr(price), _ptr(S), _ptr(K), _ptr(T), _ptr(r), _ptr(is_call), _ptr(sigma),
                                          _ptr(x_ladder), _ptr(flags[0]), _ptr(flags[1]), _ptr(ws), ws.numel(), _stream()), "ivx_bs_implied_vol_ladder")
    _count("bs_implied_vol_ladder")
    return sigma


def bs_implied_vol_chain(price, is_call, S_t, K_ek, T_te, r_t, *, n_t: int, n_exp: int, n_strike: int,
                         out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
                         want_iters: bool = False):
    """Chain layout: ``price`` per quote [t][e][k]; ``S_t``, ``r_t`` per stamp, ``K_ek`` per contract, ``T_te`` per
    (stamp, expiry); ``is_call`` per contract [e][k] (8 B/quote read) or per quote (9 B/quote) instead of 41."""
    n = n_t * n_exp * n_strike
    ticks = price.dtype == torch.int32             # int32 ticks of 1e-4 (DECIMAL(15,4)), INT32_MIN = no quote
    if not ticks:
        _req(price, torch.float64, "price")
    _req(is_call, torch.uint8, "is_call")
    per_quote = is_call.numel() == n and n != n_exp * n_strike
    if not per_quote and is_call.numel() != n_exp * n_strike:
        raise _lib.IvxError(f"is_call: expected {n_exp * n_strike} (per contract) or {n} (per quote) elements")
    for t, m, name in ((price, n, "price"), (S_t, n_t, "S_t"), (r_t, n_t, "r_t"),
                       (K_ek, n_exp * n_strike, "K_ek"), (T_te, n_t * n_exp, "T_te")):
        if t.numel() != m:
            raise _lib.IvxError(f"{name}: expected {m} elements, got {t.numel()}")
    for t in (S_t, K_ek, T_te, r_t):
        _req(t, torch.float64, "chain column")
    L = lib()
    wsb = L.ivx_bs_chain_workspace_bytes(n_t, n_exp, n_strike)
    ws = workspace if workspace is not None and workspace.numel() >= wsb else _empty(wsb, torch.uint8, price)
    sigma = out if out is not None else _empty(n, torch.float64, price)
    iters = _empty(n, torch.int32, price) if want_iters else None
    fn = L.ivx_bs_implied_vol_chain_ticks if ticks else L.ivx_bs_implied_vol_chain
    check(fn(n_t, n_exp, n_strike, _ptr(price), _ptr(is_call), int(per_quote), _ptr(S_t), _ptr(K_ek), _ptr(T_te),
                                     _ptr(r_t), _ptr(sigma), _ptr(iters), _ptr(ws), ws.numel(), _stream()),
          "ivx_bs_implied_vol_chain")
    _count("bs_implied_vol_chain")
    return (sigma, iters) if want_iters else sigma


def bs_price(S, K, T, r, sigma, is_call):
    n = S.numel()
    price = _empty(n, torch.float64, S)
    check(lib().ivx_bs_price(n, _ptr(S), _ptr(K), _ptr(T), _ptr(r), _ptr(sigma), _ptr(_req(is_call, torch.uint8, "is_call")),
                             _ptr(price), _stream()), "ivx_bs_price")
    _count("bs_price")
    return price


def bs_greeks(S, K, T, r, sigma, is_call, reference_put_rho_sign: bool = True) -> Dict[str, torch.Tensor]:
    n = S.numel()
    out = {k: _empty(n, torch.float64, S) for k in ("delta", "gamma", "theta", "vega", "rho")}
    check(lib().ivx_bs_greeks(n, _ptr(S), _ptr(K), _ptr(T), _ptr(r), _ptr(sigma), _ptr(_req(is_call, torch.uint8, "is_call")),
                              int(reference_put_rho_sign), _ptr(out["delta"]), _ptr(out["gamma"]), _ptr(out["theta"]),
                              _ptr(out["vega"]), _ptr(out["rho"]), _stream()), "ivx_bs_greeks")
    _count("bs_greeks")
    return out


# ------------------------------------------------------------------------------------------------
# Kernel D
# ------------------------------------------------------------------------------------------------
def spline_fit_eval(knot_off: torch.Tensor, x: torch.Tensor, y: torch.Tensor, xq: torch.Tensor, *,
                    max_knots: int, extrap: int = 1, out: Optional[torch.Tensor] = None,
                    workspace: Optional[torch.Tensor] = None, x_off: Optional[torch.Tensor] = None,
                    lu: Optional[Sequence[torch.Tensor]] = None, slice_ladder: Optional[torch.Tensor] = None,
                    x_own: Optional[torch.Tensor] = None, ladder_off: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Natural cubic spline per slice, evaluated on the shared grid ``xq``.  ``x_off`` (one int64 per slice) reads the
    abscissae from shared ladders; ``lu`` = the three factor arrays of ``spline_ladder_factor`` additionally lets
    slices without missing knots skip the x-only half of the elimination.  With ``x_own`` (+ ``ladder_off``) the slices
    bring their own abscissae and are only presumed to sit on their ladders: the kernel checks, bit for bit, and falls
    back to the general solve per slice."""
    n_slices = knot_off.numel() - 1
    n_q = xq.numel()
    _req(knot_off, torch.int64, "knot_off")
    for t in (x, y, xq):
        _req(t, torch.float64, "spline column")
    yq = out if out is not None else _empty(n_slices * n_q, torch.float64, x)
    if yq.numel() < n_slices * n_q:
        raise _lib.IvxError("spline_fit_eval: output too small")
    if x_off is None:
        if x.numel() != y.numel():
            raise _lib.IvxError("spline_fit_eval: x/y length mismatch")
    else:
        _req(x_off, torch.int64, "x_off")
        if x_off.numel() < n_slices:
            raise _lib.IvxError("spline_fit_eval: x_off must have one entry per slice")
    if lu is not None:
        if x_off is None or slice_ladder is None or len(lu) != 3 or any(t.dtype != torch.float64 for t in lu):
            raise _lib.IvxError("spline_fit_eval: lu needs x_off, slice_ladder and the three float64 tables of spline_ladder_factor")
        _req(slice_ladder, torch.int32, "slice_ladder")
        if x_own is not None:
            if ladder_off is None or x_own.numel() != y.numel():
                raise _lib.IvxError("spline_fit_eval: x_own needs ladder_off and one abscissa per knot")
            check(lib().ivx_spline_fit_eval_ladder_verify(n_slices, _ptr(knot_off), _ptr(x_off), _ptr(slice_ladder),
                                                          _ptr(_req(ladder_off, torch.int64, "ladder_off")), int(max_knots), _ptr(x),
                                                          _ptr(_req(x_own, torch.float64, "x_own")), _ptr(lu[0]), _ptr(lu[1]), _ptr(lu[2]),
                                                          _ptr(y), n_q, _ptr(xq), int(extrap), _ptr(yq), _stream()),
                  "ivx_spline_fit_eval_ladder_verify")
            _count("spline_fit_eval")
            return yq
        check(lib().ivx_spline_fit_eval_ladder(n_slices, _ptr(knot_off), _ptr(x_off), _ptr(slice_ladder), int(max_knots), _ptr(x),
                                               _ptr(lu[0]), _ptr(lu[1]), _ptr(lu[2]), _ptr(y), n_q, _ptr(xq), int(extrap),
                                               _ptr(yq), _stream()), "ivx_spline_fit_eval_ladder")
    else:
        check(lib().ivx_spline_fit_eval_xoff(n_slices, _ptr(knot_off), _ptr(x_off), int(max_knots), _ptr(x), _ptr(y), n_q,
                                             _ptr(xq), int(extrap), _ptr(yq), _stream()), "ivx_spline_fit_eval_xoff")
    _count("spline_fit_eval")
    return yq


def spline_ladder_factor(ladder_off: torch.Tensor, x: torch.Tensor, *, max_knots: int,
                         out: Optional[Sequence[torch.Tensor]] = None):
    """Thomas factors (c', 1/den, 1/h) of the natural-spline systems of the ladders x[ladder_off[l]:ladder_off[l+1]]: three
    tables in the order the evaluation kernel reads them (one tiny launch; reusable by every slice on those ladders; pass
    the same ``max_knots`` to ``spline_fit_eval``)."""
    _req(ladder_off, torch.int64, "ladder_off"); _req(x, torch.float64, "x")
    n_ladders = ladder_off.numel() - 1
    n_tab = lib().ivx_spline_ladder_table_len(n_ladders, int(max_knots))
    if n_tab < 0:
        raise _lib.IvxError("spline_ladder_factor: bad sizes")
    lu = out if out is not None else [torch.zeros(n_tab, dtype=torch.float64, device=x.device) for _ in range(3)]
    if any(t.numel() < n_tab for t in lu):
        raise _lib.IvxError(f"spline_ladder_factor: tables need {n_tab} elements")
    check(lib().ivx_spline_ladder_factor(n_ladders, _ptr(ladder_off), int(max_knots), _ptr(x), _ptr(lu[0]), _ptr(lu[1]),
                                         _ptr(lu[2]), _stream()), "ivx_spline_ladder_factor")
    _count("spline_ladder_factor")
    return lu


@dataclass
class SplineBand:
    """Banded evaluation operators of a set of ladders on one grid (``spline_band_build``) + the slices grouped by ladder."""
    k0: torch.Tensor                # int32[n_ladders * n_q]
    w: torch.Tensor                 # float64[n_ladders * n_q * 48]
    ok: torch.Tensor                # uint8[n_ladders]
    perm: torch.Tensor              # int16[n_ladders * 512]: grid point of every thread of the evaluation kernel
    n_q: int
    extrap: int
    frag: Optional[torch.Tensor] = None   # float64: the weights as fp64 mma B fragments per group of 8 grid points (tensor-path form)
    oct: Optional[torch.Tensor] = None    # int32: first knot and step count per group, then {largest slab extent, largest step count} per ladder
    slab_knots: int = 0                   # largest number of knots a 64-point slab touches (read back once per build; 0 = unknown)
    all_tensor_path: bool = False         # every ladder has usable tensor-path tables (read back with slab_knots)
    fill_w: Optional[torch.Tensor] = None   # float64: per ladder knot, the 48 weights that fill it when its quote is missing
    fill_k0: Optional[torch.Tensor] = None  # int32: first knot of that row (-1: the knot is not filled)


def spline_band_build(ladder_off: torch.Tensor, x_ladder: torch.Tensor, xq: torch.Tensor, *, extrap: int = 1,
                      out: Optional[SplineBand] = None, tensor_path: bool = True, refresh_hint: bool = False,
                      n_ladder_knots: Optional[int] = None) -> SplineBand:
    """Per (ladder, grid point): the 48 knot weights of "fit the natural spline, evaluate at xq[q]" and their start, checked
    against a full solve of two probe vectors (``ok[l] == 0``: that ladder's slices take the exact solve)."""
    _req(ladder_off, torch.int64, "ladder_off"); _req(x_ladder, torch.float64, "x_ladder"); _req(xq, torch.float64, "xq")
    n_l, n_q = ladder_off.numel() - 1, xq.numel()
    if out is None or out.n_q != n_q or out.ok.numel() != n_l:
        out = SplineBand(_empty(n_l * n_q, torch.int32, xq), _empty(lib().ivx_spline_band_len(n_l, n_q), torch.float64, xq),
                         _empty(n_l, torch.uint8, xq), _empty(n_l * 512, torch.int16, xq), n_q, int(extrap))
    if tensor_path and out.frag is None:
        out.frag = _empty(lib().ivx_spline_band_frag_len(n_l), torch.float64, xq)
        out.oct = _empty(lib().ivx_spline_band_oct_len(n_l), torch.int32, xq)
        n_knots_total = int(n_ladder_knots) if n_ladder_knots is not None else int(ladder_off[-1].item())   # (a read-back, once per band object)
        out.fill_w = _empty(lib().ivx_spline_band_fill_len(n_knots_total), torch.float64, xq)
        out.fill_k0 = _empty(max(1, n_knots_total), torch.int32, xq)
    out.extrap = int(extrap)
    check(lib().ivx_spline_band_build(n_l, _ptr(ladder_off), _ptr(x_ladder), n_q, _ptr(xq), int(extrap), _ptr(out.k0), _ptr(out.w),
                                      _ptr(out.ok), _ptr(out.perm), _ptr(out.frag) if tensor_path else None,
                                      _ptr(out.oct) if tensor_path else None, _ptr(out.fill_w) if tensor_path else None,
                                      _ptr(out.fill_k0) if tensor_path else None, _stream()), "ivx_spline_band_build")
    _count("spline_band_build")
    if tensor_path and n_l > 0 and (out.slab_knots == 0 or refresh_hint):
        # ONE read-back, at the first build of this set of ladders: how many knots a 64-point slab's staged rows must hold.  Later
        # builds into the same object (a column rewritten in place, a streaming pipeline's next chunk) keep the figure and stay
        # asynchronous; should a slab then need more, the kernel sees it and hands the batch to the exact solve — slower, never wrong.
        out.slab_knots = int(out.oct[n_l * 128:].view(n_l, 2)[:, 0].max().item())
        out.all_tensor_path = bool(((out.ok & 2) != 0).all().item())
    return out


def spline_fit_eval_banded(knot_off: torch.Tensor, x_off: torch.Tensor, slice_ladder: torch.Tensor, order: torch.Tensor,
                           lad_slice_off: torch.Tensor, max_slices_per_ladder: int, ladder_off: torch.Tensor, x_ladder: torch.Tensor,
                           lu: Sequence[torch.Tensor], y: torch.Tensor, xq: torch.Tensor, band: SplineBand, *, max_knots: int,
                           x_own: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                           fb: Optional[Sequence[torch.Tensor]] = None, tensor_path: bool = True, x_verified: bool = False,
                           fill_missing: bool = True) -> torch.Tensor:
    """Slices on tabulated ladders through the banded operators; whatever cannot be taken that way (a non-finite knot, a slice
    off its ladder when ``x_own`` is given, a ladder without a usable band) goes through the exact scan right behind.
    ``fb`` = (int32[n_slices] list, uint32[1] count) scratch, allocated when not passed.  ``x_verified``: ``fb`` already holds the
    verdict of comparing every slice's abscissae with its ladder's (``bs_implied_vol_ladder``); the band then does not read ``x_own``."""
    n_slices, n_q, n_l = knot_off.numel() - 1, xq.numel(), ladder_off.numel() - 1
    if x_verified and (fb is None or x_own is None):
        raise _lib.IvxError("spline_fit_eval_banded: x_verified needs the flags of the comparison (fb) and x_own for the flagged slices")
    if band.n_q != n_q:
        raise _lib.IvxError("spline_fit_eval_banded: the band was built for another grid")
    _req(order, torch.int32, "order"); _req(slice_ladder, torch.int32, "slice_ladder")
    for t_ in (knot_off, x_off, lad_slice_off, ladder_off):
        _req(t_, torch.int64, "offsets")
    yq = out if out is not None else _empty(n_slices * n_q, torch.float64, y)
    if yq.numel() < n_slices * n_q:
        raise _lib.IvxError("spline_fit_eval_banded: output too small")
    if fb is None:
        fb = (_empty(max(1, n_slices), torch.int32, y), _empty(1, torch.int32, y))
    use_mma = tensor_path and band.frag is not None
    check(lib().ivx_spline_fit_eval_banded(n_slices, _ptr(knot_off), _ptr(x_off), _ptr(slice_ladder), _ptr(order), _ptr(lad_slice_off), n_l,
                                           int(max_slices_per_ladder), _ptr(ladder_off), int(max_knots), _ptr(x_ladder),
                                           _ptr(x_own) if x_own is not None else None, _ptr(lu[0]), _ptr(lu[1]), _ptr(lu[2]), _ptr(y), n_q,
                                           _ptr(xq), int(band.extrap), _ptr(band.k0), _ptr(band.w), _ptr(band.ok), _ptr(band.perm),
                                           _ptr(band.frag) if use_mma else None, _ptr(band.oct) if use_mma else None,
                                           _ptr(band.fill_w) if (use_mma and fill_missing) else None, _ptr(band.fill_k0) if (use_mma and fill_missing) else None,
                                           int(band.slab_knots),
                                           1 if (use_mma and band.all_tensor_path) else 0, 1 if x_verified else 0, _ptr(yq), _ptr(fb[0]), _ptr(fb[1]), _stream()), "ivx_spline_fit_eval_banded")
    _count("spline_fit_eval_banded_mma" if use_mma else "spline_fit_eval_banded")
    return yq


def spline_strided(x: torch.Tensor, y: torch.Tensor, xq: torch.Tensor, n_systems: int, *, extrap: int = 1,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    n_knots, n_q = x.numel(), xq.numel()
    for t in (x, y, xq):
        _req(t, torch.float64, "spline column")
    if y.numel() != n_knots * n_systems:
        raise _lib.IvxError("y must be [n_knots][n_systems]")
    yq = out if out is not None else _empty(n_q * n_systems, torch.float64, x)
    check(lib().ivx_spline_strided(n_systems, n_knots, _ptr(x), _ptr(y), n_q, _ptr(xq), int(extrap), _ptr(yq),
                                   _stream()), "ivx_spline_strided")
    _count("spline_strided")
    return yq


def spline_strided_batched(x: torch.Tensor, x_batch_stride: int, x_knot_stride: int, y: torch.Tensor, xq: torch.Tensor,
                           n_batches: int, n_systems: int, n_knots: int, *, extrap: int = 1,
                           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``n_batches`` planes of ``n_systems`` splines: y[b][knot][system] -> yq[b][query][system]; the abscissae of
    plane b are read in place at ``x[b*x_batch_stride + i*x_knot_stride]``."""
    n_q = xq.numel()
    for t in (x, y, xq):
        _req(t, torch.float64, "spline column")
    if y.numel() != n_batches * n_knots * n_systems:
        raise _lib.IvxError("y must be [n_batches][n_knots][n_systems]")
    if n_batches and x.numel() <= (n_batches - 1) * x_batch_stride + (n_knots - 1) * x_knot_stride:
        raise _lib.IvxError("x is too short for the given strides")
    yq = out if out is not None else _empty(n_batches * n_q * n_systems, torch.float64, y)
    check(lib().ivx_spline_strided_batched(n_batches, n_systems, n_knots, _ptr(x), int(x_batch_stride), int(x_knot_stride),
                                           _ptr(y), n_q, _ptr(xq), int(extrap), _ptr(yq), _stream()),
          "ivx_spline_strided_batched")
    _LAUNCH_EXTRA = (n_batches + 65534) // 65535
    global _LAUNCHES
    _LAUNCHES += max(1, _LAUNCH_EXTRA)
    return yq


# ------------------------------------------------------------------------------------------------
# Data bridge
# ------------------------------------------------------------------------------------------------
@dataclass
class BridgePlan:
    out_off: torch.Tensor          # int64[n_sym+1] emitted candles before each symbol
    e_off: torch.Tensor            # int64[n_sym+1] exponential draws before each symbol
    seg_out: torch.Tensor          # int64[n_seg+1] the same before each 2048-row segment
    seg_e: torch.Tensor
    total: int
    total_e: int


@dataclass
class BridgeBatch:
    out_off: torch.Tensor          # int64[n_sym+1]
    row: torch.Tensor              # int64[m] global input row of each candle
    open: torch.Tensor
    high: torch.Tensor
    low: torch.Tensor
    close: torch.Tensor
    volume: torch.Tensor
    source_price: torch.Tensor
    sym_quality: torch.Tensor      # int32[n_sym] bit flags of _validate_ohlcv_quality
    total: int
    ts: Optional[torch.Tensor] = None      # int64[m] stamp of each candle (bridge_ohlcv_philox with ts=; `row` may then be None)


def bridge_plan(in_off: torch.Tensor, base: torch.Tensor, volume: torch.Tensor) -> BridgePlan:
    """Rows that emit a candle (base price neither NaN nor <= 0, ohlcv_converter.py:152) and rows among them whose
    volume is drawn (NaN or <= 0, :359-369), counted before every symbol and before every 2048-row segment: the layout of
    the outputs and of the tapes of ``bridge_ohlcv``."""
    n_sym, n_rows = in_off.numel() - 1, base.numel()
    n_seg = lib().ivx_bridge_n_segments(n_rows)
    seg_out, seg_e = _empty(n_seg + 1, torch.int64, in_off), _empty(n_seg + 1, torch.int64, in_off)
    out_off, e_off = _empty(n_sym + 1, torch.int64, in_off), _empty(n_sym + 1, torch.int64, in_off)
    wsb = lib().ivx_bridge_workspace_bytes(n_rows)
    ws = torch.empty(wsb, dtype=torch.uint8, device=in_off.device)
    check(lib().ivx_bridge_plan(n_sym, n_rows, _ptr(_req(in_off, torch.int64, "in_off")), _ptr(_req(base, torch.float64, "base")),
                                _ptr(_req(volume, torch.float64, "volume")), _ptr(seg_out), _ptr(seg_e), _ptr(out_off), _ptr(e_off),
                                _ptr(ws), wsb, _stream()), "ivx_bridge_plan")
    _count("bridge_plan")
    total, total_e = read_back(out_off[-1:])[0], read_back(e_off[-1:])[0]          # sizes of the outputs and of the tapes
    return BridgePlan(out_off, e_off, seg_out, seg_e, total, total_e)


def bridge_ohlcv(in_off: torch.Tensor, base: torch.Tensor, volume: torch.Tensor, plan: BridgePlan, strategy: str, *,
                 u: Optional[torch.Tensor] = None, e: Optional[torch.Tensor] = None, z: Optional[torch.Tensor] = None,
                 base_spread_percent: float = 0.002, volatility_factor: float = 1.5, segmented: bool = True,
                 recompute: bool = False) -> BridgeBatch:
    """Synthetic 1-minute candles from interpolated rows, every candle bit-equal to the reference's given the same
    random tapes (see include/ivx.h).  ``strategy`` is one of _lib.BRIDGE_STRATEGY.  ``segmented=False`` runs one warp per
    symbol instead of one per 2048-row segment (same results)."""
    code = _lib.BRIDGE_STRATEGY[strategy]
    k = _lib.BRIDGE_U_PER_ROW[code]
    n_sym = in_off.numel() - 1
    m = plan.total
    if k and (u is None or u.numel() < k * m):
        raise _lib.IvxError(f"bridge_ohlcv: strategy {strategy} needs {k * m} uniforms, got {0 if u is None else u.numel()}")
    if code == 2 and (z is None or z.numel() < m):
        raise _lib.IvxError(f"bridge_ohlcv: trend_following needs {m} normals, got {0 if z is None else z.numel()}")
    if e is None or e.numel() < plan.total_e:
        raise _lib.IvxError(f"bridge_ohlcv: {plan.total_e} exponential draws needed, got {0 if e is None else e.numel()}")
    if e.numel() == 0:
        e = _empty(1, torch.float64, base)
    cols = [_empty(m, torch.float64, base) for _ in range(6)]
    row = _empty(m, torch.int64, base)
    q = torch.zeros(n_sym, dtype=torch.int32, device=base.device)
    use_seg = segmented and code != 2
    ws, wsb = None, 0
    if use_seg:
        wsb = lib().ivx_bridge_workspace_bytes(base.numel())
        ws = torch.empty(wsb, dtype=torch.uint8, device=base.device)
    check(lib().ivx_bridge_ohlcv(n_sym, base.numel(), _ptr(in_off), _ptr(_req(base, torch.float64, "base")), _ptr(_req(volume, torch.float64, "volume")),
                                 code | (_lib.BRIDGE_RECOMPUTE if recompute else 0), float(base_spread_percent), float(volatility_factor),
                                 _ptr(_req(u, torch.float64, "u")) if u is not None else None, _ptr(_req(e, torch.float64, "e")),
                                 _ptr(_req(z, torch.float64, "z")) if z is not None else None,
                                 _ptr(plan.out_off), _ptr(plan.e_off), _ptr(plan.seg_out) if use_seg else None,
                                 _ptr(plan.seg_e) if use_seg else None, _ptr(row), *(_ptr(c) for c in cols), _ptr(q),
                                 _ptr(ws), wsb, _stream()),
          "ivx_bridge_ohlcv")
    _count("bridge_ohlcv_seg" if use_seg else "bridge_ohlcv")
    return BridgeBatch(plan.out_off, row, *cols, q, m)


def bridge_ohlcv_philox(in_off: torch.Tensor, base: torch.Tensor, volume: torch.Tensor, plan: BridgePlan, strategy: str, *,
                        seed: int, stream_u: int, stream_e: int, ts: Optional[torch.Tensor] = None, want_row: bool = True,
                        base_spread_percent: float = 0.002, volatility_factor: float = 1.5, segmented: bool = True,
                        recompute: bool = False) -> BridgeBatch:
    """``bridge_ohlcv`` with the tapes generated inside the kernel: the candles are bit-equal to
    ``bridge_ohlcv(u=random_tape(k m, seed, stream_u), e=random_tape(m_e, seed, stream_e, "exponential"))`` without the tapes ever
    being written or read.  ``ts`` (int64 per input row): the candles' stamps come back in ``.ts`` (no gather behind the
    bridge); ``want_row=False`` then leaves ``.row`` out."""
    code = _lib.BRIDGE_STRATEGY[strategy]
    if code == 2:
        raise _lib.IvxError("bridge_ohlcv_philox: trend_following draws normals (pass a tape to bridge_ohlcv)")
    if ts is None and not want_row:
        raise _lib.IvxError("bridge_ohlcv_philox: want_row=False needs ts")
    n_sym, m = in_off.numel() - 1, plan.total
    cols = [_empty(m, torch.float64, base) for _ in range(6)]
    row = _empty(m, torch.int64, base) if want_row else None
    o_ts = _empty(m, torch.int64, base) if ts is not None else None
    q = torch.zeros(n_sym, dtype=torch.int32, device=base.device)
    ws, wsb = None, 0
    if segmented:
        wsb = lib().ivx_bridge_workspace_bytes(base.numel())
        ws = torch.empty(wsb, dtype=torch.uint8, device=base.device)
    check(lib().ivx_bridge_ohlcv_philox(n_sym, base.numel(), _ptr(in_off), _ptr(_req(base, torch.float64, "base")),
                                        _ptr(_req(volume, torch.float64, "volume")), _ptr(_req(ts, torch.int64, "ts")) if ts is not None else None,
                                        code | (_lib.BRIDGE_RECOMPUTE if recompute else 0), float(base_spread_percent), float(volatility_factor),
                                        int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_u) & 0xFFFFFFFF, int(stream_e) & 0xFFFFFFFF,
                                        _ptr(plan.out_off), _ptr(plan.e_off), _ptr(plan.seg_out) if segmented else None,
                                        _ptr(plan.seg_e) if segmented else None, _ptr(row), _ptr(o_ts), *(_ptr(c) for c in cols), _ptr(q),
                                        _ptr(ws), wsb, _stream()),
          "ivx_bridge_ohlcv_philox")
    _count("bridge_ohlcv_seg" if segmented else "bridge_ohlcv")
    return BridgeBatch(plan.out_off, row, *cols, q, m, o_ts)


def bridge_quality(o: torch.Tensor, h: torch.Tensor, l: torch.Tensor, c: torch.Tensor, src: torch.Tensor) -> int:
    """Flag bits of _validate_ohlcv_quality (ohlcv_converter.py:371-402) over existing candles."""
    flags = torch.zeros(1, dtype=torch.int32, device=o.device)
    check(lib().ivx_bridge_quality(o.numel(), *(_ptr(_req(t, torch.float64, "ohlc")) for t in (o, h, l, c, src)), _ptr(flags), _stream()),
          "ivx_bridge_quality")
    _count("bridge_quality")
    return int(flags.item())


# ------------------------------------------------------------------------------------------------
# Utilities
# ------------------------------------------------------------------------------------------------
def exclusive_scan(x: torch.Tensor) -> torch.Tensor:
    out = _empty(x.numel() + 1, torch.int64, x)
    check(lib().ivx_exclusive_scan_i64(x.numel(), _ptr(_req(x, torch.int64, "x")), _ptr(out), _stream()), "ivx_exclusive_scan_i64")
    _count("scan")
    return out


def gather(idx: torch.Tensor, src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``src[idx]`` for int64 or float64 ``src`` (NaT / NaN where ``idx`` is out of range)."""
    _req(idx, torch.int64, "idx")
    n = idx.numel()
    out = out if out is not None else _empty(n, src.dtype, src)
    if out.numel() < n or out.dtype != src.dtype:
        raise _lib.IvxError("gather: output too small or of the wrong dtype")
    if src.dtype == torch.int64:
        fn = lib().ivx_gather_i64
    elif src.dtype == torch.float64:
        fn = lib().ivx_gather_f64
    else:
        raise _lib.IvxError(f"gather: int64 or float64 source, got {src.dtype}")
    check(fn(n, _ptr(idx), src.numel(), _ptr(src), _ptr(out), _stream()), "ivx_gather")
    _count("gather")
    return out[:n]


def random_tape(n: int, seed: int, stream_id: int = 0, kind: str = "uniform", like: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``n`` uniforms in [0, 1) or standard exponentials from the device's counter generator (element i depends on
    (seed, stream_id, i) only): the bridge's tapes without a host draw."""
    if out is None:
        if like is None:
            raise _lib.IvxError("random_tape: pass `like` (a CUDA tensor) or `out`")
        out = _empty(max(1, n), torch.float64, like)
    _req(out, torch.float64, "out")
    if out.numel() < n:
        raise _lib.IvxError("random_tape: output too small")
    check(lib().ivx_random_tape(int(n), int(seed) & (2 ** 64 - 1), int(stream_id) & 0xffffffff, {"uniform": 0, "exponential": 1}[kind],
                                _ptr(out), _stream()), "ivx_random_tape")
    _count("random_tape")
    return out


def probe_copy(src: torch.Tensor, dst: torch.Tensor) -> None:
    check(lib().ivx_probe_copy(src.numel(), _ptr(src), _ptr(dst), _stream()), "ivx_probe_copy")
    _count("probe")


def probe_fp64_fma(n_threads: int, iters: int, sink: torch.Tensor) -> None:
    check(lib().ivx_probe_fp64_fma(n_threads, iters, _ptr(sink), _stream()), "ivx_probe_fp64_fma")
    _count("probe")
