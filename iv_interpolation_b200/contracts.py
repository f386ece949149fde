"""The reference's three-stage pipeline for a CSR batch of contracts, resident on one GPU or streamed from host memory.

    Task 1   IVInterpolator.interpolate_symbol              reference src/interpolation/core.py:16-85       (kernel A)
    bridge   InterpolatedToOHLCVConverter._generate_...     reference src/data_bridge/ohlcv_converter.py:138-263
    Task 2   CandleReconstructor.reconstruct_symbol_candles reference src/candle_reconstruction/core.py:36-106 (kernel B)

— the order ``main.py --task pipeline`` / ``complete_pipeline.py`` run them in (reference complete_pipeline.py:740-831),
one launch per stage for ALL contracts of the batch instead of one pandas call chain per contract.  Contracts are
independent (the reference fetches ``WHERE symbol = %s``, batch_processor.py:81): a shard of contracts per GPU, no
exchange between GPUs; the optional last step is the gather of the 5-minute candle tables (``sharding.gather_tables``).

``ContractPipeline.run`` takes device columns; ``HostContractStream.run`` takes pinned host columns, copies them in
batch by batch on a copy stream while the previous batch computes, and brings each batch's 5-minute candles back to
pinned host memory — the end-to-end entry ``bench.py --workload contracts`` times.  All arithmetic is in libivx.so.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import device as D
from . import tapes as TP
from ._lib import BRIDGE_STRATEGY, BRIDGE_U_PER_ROW, IvxError

STAGES = ("interp_plan", "interp_exec", "bridge_plan", "bridge_tapes", "bridge_ohlcv", "candle_ts", "ohlcv_reduce")
SRC_COLS = TP.INTERP_COLS + TP.FILL_NUM_COLS            # the nine numeric source columns, in launch order


@dataclass
class ContractResult:
    interp: D.InterpBatch          # Task 1 rows (contract-major minute rows)
    minute: D.BridgeBatch          # synthetic 1-minute candles (CSR by contract: minute.out_off)
    minute_ts: torch.Tensor        # int64[minute.total]
    candles: D.CandleBatch         # N-minute candles (CSR by contract: candles.out_off)

    def candle_table(self, first_contract: int = 0) -> Dict[str, torch.Tensor]:
        """The 5-minute candles as flat columns + the (shard-local) contract each belongs to (for the final gather)."""
        c = self.candles
        n_sym = c.out_off.numel() - 1
        # output_size: without it torch reads the total back with a cudaMemcpy, which queues behind every bulk device-to-host copy
        # in flight (HostContractStream calls this right after it has enqueued the batch's drain: the host then waited for the
        # whole drain, and compute and drain stopped overlapping)
        sym = torch.repeat_interleave(torch.arange(first_contract, first_contract + n_sym, dtype=torch.int32, device=c.ts.device),
                                      (c.out_off[1:] - c.out_off[:-1]), output_size=int(c.total))
        return {"contract": sym, "timestamp": c.ts, "open": c.open, "high": c.high, "low": c.low, "close": c.close,
                "volume": c.volume, "count": c.count}


class ContractPipeline:
    """Task 1 -> bridge -> Task 2 on device-resident columns."""

    def __init__(self, device: torch.device, *, strategy: str = "spread_simulation", frequency_minutes: int = 5,
                 seed: int = 0, step: int = TP.MINUTE_US, max_span: int = 30 * TP.DAY_US, min_points: int = 10,
                 base_spread_percent: float = 0.002, volatility_factor: float = 1.5, fused_tapes: bool = True,
                 want_row: bool = True):
        if strategy not in BRIDGE_STRATEGY or strategy == "trend_following":
            raise IvxError(f"ContractPipeline: strategy {strategy!r} (trend_following takes normals: use device.bridge_ohlcv)")
        self.device = device
        self.strategy = strategy
        self.n = int(frequency_minutes)
        self.seed = int(seed)
        self.step, self.max_span, self.min_points = int(step), int(max_span), int(min_points)
        self.spread = dict(base_spread_percent=base_spread_percent, volatility_factor=volatility_factor) \
            if strategy == "spread_simulation" else {}
        self.fused_tapes = fused_tapes                   # False: tapes written by random_tape, read by the bridge, stamps gathered (A/B)
        self.want_row = want_row                         # False: the candles keep their stamps only (minute.row is None)
        self.calls = 0

    def run(self, off: torch.Tensor, date: torch.Tensor, icols: Sequence[torch.Tensor], fcols: Sequence[torch.Tensor],
            *, tapes: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, events: Optional[dict] = None,
            first_contract: int = 0) -> ContractResult:
        """``fcols`` in ``tapes.FILL_NUM_COLS`` order (volume is what the bridge reads).  ``tapes=(u, e)``: host-drawn
        random tapes in the reference's order (parity runs); default: the device's counter generator, keyed by
        (seed, first_contract) so that a shard draws the same numbers wherever it is processed."""
        def mark(stage, k):
            if events is not None:
                events[stage][k].record()

        mark("interp_plan", 0)
        plan = D.interp_plan(off, date, icols, fcols, [None], step=self.step, max_span=self.max_span, min_points=self.min_points)
        plan.total                                        # 8-byte read-back: sizes the outputs
        mark("interp_plan", 1)
        mark("interp_exec", 0)
        a = D.interp_exec(plan, want_is_interpolated=True)
        mark("interp_exec", 1)
        S, vol = a.interp[1], a.fill[TP.FILL_NUM_COLS.index("volume")]
        mark("bridge_plan", 0)
        bplan = D.bridge_plan(a.out_off, S, vol)
        mark("bridge_plan", 1)
        if tapes is None and self.fused_tapes:
            # the reference's normal, unseeded mode: the draws are made inside the bridge kernel (Philox keyed by the seed and the
            # shard's first contract, the numbers random_tape would have written) and the candle stamps are copied there too —
            # no tape is written or read and no gather pass follows; the two stages below stay in the table at zero
            mark("bridge_tapes", 0); mark("bridge_tapes", 1)
            mark("bridge_ohlcv", 0)
            m = D.bridge_ohlcv_philox(a.out_off, S, vol, bplan, self.strategy, seed=self.seed, stream_u=2 * first_contract,
                                      stream_e=2 * first_contract + 1, ts=a.ts, want_row=self.want_row, **self.spread)
            mark("bridge_ohlcv", 1)
            mark("candle_ts", 0); mark("candle_ts", 1)
            mts = m.ts
        else:
            mark("bridge_tapes", 0)
            if tapes is None:
                k = BRIDGE_U_PER_ROW[BRIDGE_STRATEGY[self.strategy]]
                u = D.random_tape(k * bplan.total, self.seed, 2 * first_contract, "uniform", like=S)
                e = D.random_tape(bplan.total_e, self.seed, 2 * first_contract + 1, "exponential", like=S)
            else:
                u, e = tapes
            mark("bridge_tapes", 1)
            mark("bridge_ohlcv", 0)
            m = D.bridge_ohlcv(a.out_off, S, vol, bplan, self.strategy, u=u, e=e, **self.spread)
            mark("bridge_ohlcv", 1)
            mark("candle_ts", 0)
            mts = D.gather(m.row, a.ts)
            mark("candle_ts", 1)
        mark("ohlcv_reduce", 0)
        c = D.ohlcv_reduce(m.out_off, mts, m.open, m.high, m.low, m.close, m.volume,
                           bucket_ticks=self.n * self.step, min_count=self.n, check_len=True)
        mark("ohlcv_reduce", 1)
        self.calls += 1
        return ContractResult(a, m, mts, c)


# ----------------------------------------------------------------------------------------------------------
# host side: a shard of contracts in pinned memory, streamed through the pipeline
# ----------------------------------------------------------------------------------------------------------
@dataclass
class HostContractShard:
    """Source rows of a shard's contracts in pinned host memory: CSR offsets + date + the nine numeric columns
    (what reference batch_processor.py:76-83 SELECTs, minus the three string columns, which never reach the device)."""
    off: torch.Tensor                         # int64[n+1], CPU
    date: torch.Tensor                        # int64[rows], pinned
    cols: Dict[str, torch.Tensor]             # float64[rows] each, pinned
    out_rows: np.ndarray                      # minute rows each contract will produce (cost model, sharding.contract_costs)
    ids: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))   # global contract ids

    @property
    def n_contracts(self) -> int:
        return self.off.numel() - 1

    @classmethod
    def from_tape(cls, tp: "TP.ContractTape", ids: Optional[np.ndarray] = None, pin: bool = True) -> "HostContractShard":
        from .sharding import contract_costs
        f = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()) if pin else (lambda a: torch.from_numpy(np.ascontiguousarray(a)))
        return cls(torch.from_numpy(np.ascontiguousarray(tp.off)), f(tp.date), {c: f(tp.cols[c]) for c in SRC_COLS},
                   contract_costs(tp.off, tp.date, TP.MINUTE_US), np.arange(tp.n_sym) if ids is None else np.asarray(ids))

    @classmethod
    def concat(cls, shards: Sequence["HostContractShard"], pin: bool = True) -> "HostContractShard":
        offs, base = [np.zeros(1, np.int64)], 0
        for s in shards:
            o = s.off.numpy()
            offs.append(o[1:] + base)
            base += int(o[-1])
        cat = lambda ts: (torch.cat(list(ts)).pin_memory() if pin else torch.cat(list(ts)))
        return cls(torch.from_numpy(np.concatenate(offs)), cat(s.date for s in shards),
                   {c: cat(s.cols[c] for s in shards) for c in SRC_COLS}, np.concatenate([s.out_rows for s in shards]),
                   np.concatenate([s.ids for s in shards]))


class HostContractStream:
    """Pinned host source rows in, 5-minute candles out (pinned host ring + resident device tables).

    Batches of contracts are cut so that a batch's minute rows stay below ``max_rows_per_batch`` (the Task-1 output and
    the bridge's tapes and candles of one batch are what has to fit in HBM at once).  Batch k+1's source rows are copied
    on the copy stream while batch k computes; batch k's candle table leaves on the drain stream."""

    CANDLE_COLS = ("timestamp", "open", "high", "low", "close", "volume", "count")

    def __init__(self, device: torch.device, *, max_rows_per_batch: int = 160_000_000, keep_device_tables: bool = True, **pipeline_kw):
        self.device = device
        self.pipe = ContractPipeline(device, **pipeline_kw)
        self.max_rows = int(max_rows_per_batch)
        self.keep = keep_device_tables
        self.copy_stream = torch.cuda.Stream(device)
        self.drain_stream = torch.cuda.Stream(device)
        self.h2d_bytes = self.d2h_bytes = 0
        self.ring: Optional[Dict[str, torch.Tensor]] = None
        self.ring_rows = 0

    def batches(self, shard: HostContractShard) -> List[Tuple[int, int]]:
        cuts, a, acc = [], 0, 0
        for i, r in enumerate(shard.out_rows):
            if acc and acc + r > self.max_rows:
                cuts.append((a, i)); a, acc = i, 0
            acc += int(r)
        if a < shard.n_contracts:
            cuts.append((a, shard.n_contracts))
        return cuts

    def _ring(self, rows: int):
        """Pinned landing area for the candle tables: a ring the consumer reads out of; sized for the largest batch."""
        if self.ring is None or self.ring_rows < rows:
            self.ring_rows = int(rows * 1.05) + 1024
            mk = lambda dt: torch.empty(self.ring_rows, dtype=dt).pin_memory()
            self.ring = {c: mk(torch.int64 if c == "timestamp" else torch.int32 if c == "count" else torch.float64)
                         for c in self.CANDLE_COLS}
            self.ring["out_off"] = torch.empty(1 << 22, dtype=torch.int64).pin_memory()
        return self.ring

    def _h2d(self, shard: HostContractShard, a: int, b: int):
        off = shard.off
        r0, r1 = int(off[a]), int(off[b])
        with torch.cuda.stream(self.copy_stream):
            d_off = (off[a:b + 1] - r0).to(self.device, non_blocking=True)
            date = shard.date[r0:r1].to(self.device, non_blocking=True)
            cols = {c: shard.cols[c][r0:r1].to(self.device, non_blocking=True) for c in SRC_COLS}
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.h2d_bytes += (b - a + 1) * 8 + (r1 - r0) * 8 * (1 + len(SRC_COLS))
        return d_off, date, cols, ev

    def run(self, shard: HostContractShard, events: Optional[list] = None) -> Dict[str, object]:
        """Returns {"tables": [per-batch device candle tables] (if kept), "candles": total, "minute_rows": total}."""
        self.h2d_bytes = self.d2h_bytes = 0
        cuts = self.batches(shard)
        main = torch.cuda.current_stream(self.device)
        tables, n_candles, n_rows, n_minute = [], 0, 0, 0
        nxt = self._h2d(shard, *cuts[0]) if cuts else None
        drained: Optional[torch.cuda.Event] = None
        for k, (a, b) in enumerate(cuts):
            d_off, date, cols, ev = nxt
            nxt = self._h2d(shard, *cuts[k + 1]) if k + 1 < len(cuts) else None
            main.wait_event(ev)
            res = self.pipe.run(d_off, date, [cols[c] for c in TP.INTERP_COLS], [cols[c] for c in TP.FILL_NUM_COLS],
                                first_contract=int(shard.ids[a]) if len(shard.ids) else a,
                                events=events[k] if events is not None else None)
            for t in (d_off, date, *cols.values()):
                t.record_stream(main)
            c = res.candles
            m = c.total
            ring = self._ring(m)
            done = torch.cuda.Event(); done.record(main)
            with torch.cuda.stream(self.drain_stream):
                self.drain_stream.wait_event(done)
                for name, t in (("timestamp", c.ts), ("open", c.open), ("high", c.high), ("low", c.low), ("close", c.close),
                                ("volume", c.volume), ("count", c.count)):
                    ring[name][:m].copy_(t, non_blocking=True)
                    t.record_stream(self.drain_stream)
                ring["out_off"][: b - a + 1].copy_(c.out_off, non_blocking=True)
                drained = torch.cuda.Event(); drained.record(self.drain_stream)
            self.d2h_bytes += m * (8 * 6 + 4) + (b - a + 1) * 8
            n_candles += m; n_rows += res.interp.total; n_minute += res.minute.total
            if self.keep:
                tables.append(res.candle_table(a))
        if drained is not None:
            main.wait_event(drained)
        return {"tables": tables, "candles": n_candles, "interp_rows": n_rows, "minute_candles": n_minute, "batches": len(cuts)}
