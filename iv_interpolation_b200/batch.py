"""DB-free batch drivers with the reference's per-symbol status / summary dictionaries.

The reference's drivers wrap the two hot-path classes in PostgreSQL I/O: fetch one symbol, call the class,
INSERT the rows, update a progress table (reference src/interpolation/batch_processor.py:67-142,202-293;
src/candle_reconstruction/core.py:207-270).  The I/O is out of scope; what callers see —
the status dictionaries — is kept, with frames in and frames out and ONE kernel launch per stage for the whole
batch instead of one pandas call chain per symbol.  ``sink(symbol, frame) -> bool`` stands where the reference
saves to the database (`_save_results`, `_save_reconstructed_candles`).
"""
from __future__ import annotations

import time
from typing import Callable, Dict, List, Mapping, Optional, Tuple

import numpy as np
import pandas as pd

from ._lib import FLAG_IN_BAD_OHLC, FLAG_IN_NAN, FLAG_IN_NEG_VOLUME, FLAG_OUT_BAD_OHLC, FLAG_OUT_NAN, FLAG_OUT_NEG_VOLUME
from .candle_reconstruction.core import CandleReconstructor
from .interpolation.core import IVInterpolator

Sink = Callable[[str, pd.DataFrame], bool]
_IN_BAD = FLAG_IN_BAD_OHLC | FLAG_IN_NEG_VOLUME | FLAG_IN_NAN
_OUT_BAD = FLAG_OUT_BAD_OHLC | FLAG_OUT_NEG_VOLUME | FLAG_OUT_NAN


def _summary(batch_id: int, statuses: List[dict], t0: float) -> dict:
    """batch_processor.py:287-293."""
    return {"batch_id": batch_id,
            "success_count": sum(s["status"] == "success" for s in statuses),
            "error_count": sum(s["status"] == "error" for s in statuses),
            "skip_count": sum(s["status"] == "skipped" for s in statuses),
            "duration": time.time() - t0}


class InterpolationBatch:
    """``BatchProcessor.process_symbol`` for every symbol of a batch (batch_processor.py:67-142)."""

    def __init__(self, method: str = "linear", min_points: int = 10, sink: Optional[Sink] = None, batch_id: Optional[int] = None,
                 recompute_greeks: bool = False, reference_put_rho_sign: bool = True):
        self.interpolator = IVInterpolator(method=method, min_points=min_points)      # batch_processor.py:21-24
        self.sink = sink
        # the reference promises Greeks on the interpolated rows (config.py:46 preserve_greeks, schema.py:36-40) but never
        # wires greeks.py in; here they are one more launch over the batch's rows
        self.recompute_greeks = recompute_greeks
        self.reference_put_rho_sign = reference_put_rho_sign
        self.batch_id = int(time.time()) if batch_id is None else batch_id            # monitoring/progress.py:18-20

    def run(self, frames: Mapping[str, pd.DataFrame], whole_out: Optional[dict] = None) -> Tuple[Dict[str, pd.DataFrame], List[dict], dict]:
        t0 = time.time()
        symbols = list(frames)
        statuses: List[Optional[dict]] = [None] * len(symbols)
        todo = []
        for i, sym in enumerate(symbols):
            if frames[sym] is None or frames[sym].empty:                              # :89-92
                statuses[i] = {"symbol": sym, "status": "skipped", "reason": "No data found"}
            else:
                todo.append(i)
        outs: Dict[str, pd.DataFrame] = {}
        try:
            results = self.interpolator.interpolate_batch([frames[symbols[i]] for i in todo], extra={"batch_id": self.batch_id},
                                                          whole_out=whole_out) if todo else []                     # :105
            if whole_out is not None and "slots" in whole_out:            # positions in `symbols` the whole frame's offsets stand for
                whole_out["names"] = [symbols[todo[k]] for k in whole_out["slots"]]
            err = None
        except Exception as e:                                                        # :132-142
            results, err = [None] * len(todo), str(e)
        if self.recompute_greeks and err is None:
            try:
                self._add_greeks([r for r in results if r is not None])
            except Exception as e:
                results, err = [None] * len(todo), str(e)
        elapsed = time.time() - t0
        for i, res in zip(todo, results):
            sym = symbols[i]
            if err is not None:
                statuses[i] = {"symbol": sym, "status": "error", "error": err, "processing_time": elapsed}
            elif res is None:                                                         # :99-102
                statuses[i] = {"symbol": sym, "status": "skipped", "reason": "Interpolation failed or insufficient data"}
            else:
                if self.sink is None or self.sink(sym, res):                          # :108
                    outs[sym] = res
                    statuses[i] = {"symbol": sym, "status": "success", "input_rows": len(frames[sym]),
                                   "output_rows": len(res), "processing_time": elapsed / max(1, len(todo))}
                else:
                    statuses[i] = {"symbol": sym, "status": "error", "error": "Failed to save to database"}
        return outs, statuses, _summary(self.batch_id, statuses, t0)


    def _add_greeks(self, frames: List[pd.DataFrame]) -> None:
        """delta, gamma, theta, vega, rho per interpolated row (reference src/interpolation/greeks.py:12-43) with
        S = underlying_price, K = strike, T = time_to_maturity, r = interest_rate, sigma = iv, type from callput."""
        import numpy as np
        import torch

        from . import device as D
        frames = [f for f in frames if len(f) and all(c in f.columns for c in ("strike", "callput", "interest_rate"))]
        if not frames:
            return
        dev = torch.device("cuda")
        cat = lambda c: torch.from_numpy(np.concatenate([f[c].to_numpy(dtype=np.float64) for f in frames])).to(dev)
        call = np.concatenate([(f["callput"].astype(str).str.lower().str[0] == "c").to_numpy() for f in frames]).astype(np.uint8)
        g = D.bs_greeks(cat("underlying_price"), cat("strike"), cat("time_to_maturity"), cat("interest_rate"), cat("iv"),
                        torch.from_numpy(call).to(dev), reference_put_rho_sign=self.reference_put_rho_sign)
        host = {k: v.cpu().numpy() for k, v in g.items()}
        a = 0
        for f in frames:
            b = a + len(f)
            for k in ("delta", "gamma", "theta", "vega", "rho"):
                f[k] = host[k][a:b]
            a = b


class CandleBatch:
    """``MultiSymbolCandleReconstructor.process_symbol`` for every symbol of a batch (candle core.py:207-270):
    validate input -> reconstruct -> validate output -> save -> stats, the two validations read from the flag
    bits the reduction kernel produces in the same launch."""

    def __init__(self, target_frequency: str = "5min", sink: Optional[Sink] = None):
        self.reconstructor = CandleReconstructor(target_frequency)                    # candle core.py:180
        self.sink = sink

    def run(self, frames: Mapping[str, pd.DataFrame]) -> Tuple[Dict[str, pd.DataFrame], List[dict], dict]:
        t0 = time.time()
        symbols = list(frames)
        statuses: List[Optional[dict]] = [None] * len(symbols)
        todo = []
        for i, sym in enumerate(symbols):
            if frames[sym] is None or frames[sym].empty:                              # :225-226
                statuses[i] = {"symbol": sym, "status": "skipped", "reason": "No minute data found"}
            else:
                todo.append(i)
        flags: list = []
        outs: Dict[str, pd.DataFrame] = {}
        try:
            # check_len=False: the input validation (:229) comes before the length early-out inside reconstruct (:64)
            results = self.reconstructor.reconstruct_batch([frames[symbols[i]] for i in todo], check_len=False,
                                                           flags_out=flags) if todo else []
            err = None
        except Exception as e:
            results, err, flags = [None] * len(todo), str(e), [None] * len(todo)
        elapsed = time.time() - t0
        n = self.reconstructor.frequency_minutes
        for k, (i, res) in enumerate(zip(todo, results)):
            sym, src = symbols[i], frames[symbols[i]]
            fl = flags[k]
            if err is not None:
                statuses[i] = {"symbol": sym, "status": "error", "error": err, "processing_time": elapsed}
            elif any(c not in src.columns for c in ("open", "high", "low", "close", "volume")) or \
                    (fl is not None and (fl & _IN_BAD)):                              # :229-230 (validate raises inside -> False)
                statuses[i] = {"symbol": sym, "status": "error", "error": "Invalid input candle data"}
            elif res is None or len(src) < n:                                         # :235-236 (None incl. :64-66)
                statuses[i] = {"symbol": sym, "status": "skipped", "reason": "Reconstruction failed"}
            elif res.empty or (fl & _OUT_BAD):                                        # :239-240: validate(empty) is False
                statuses[i] = {"symbol": sym, "status": "error", "error": "Invalid reconstructed candle data"}
            elif self.sink is not None and not self.sink(sym, res):
                statuses[i] = {"symbol": sym, "status": "error", "error": "Failed to save reconstructed candles"}
            else:
                outs[sym] = res
                statuses[i] = {"symbol": sym, "status": "success", "input_candles": len(src), "output_candles": len(res),
                               "processing_time": elapsed / max(1, len(todo)),
                               "stats": self.reconstructor.get_reconstruction_stats(src, res)}
        return outs, statuses, _summary(int(t0), statuses, t0)


class PipelineBatch:
    """``CompleteOptimizedPipeline``'s three stages for a batch of symbols, DB-free (reference complete_pipeline.py):
    Task 1 interpolation (:318), the data bridge ``_convert_symbol_to_ohlcv`` / ``_generate_ohlcv_candles`` (:437-510) and
    Task 2 ``_reconstruct_symbol_candles`` / ``_aggregate_to_5min`` (:641-710) — each stage ONE launch for all symbols, the
    reference's per-symbol status dictionaries per stage.  The bridge draws from numpy's global generator (or ``rng``) in
    the order the reference does when it loops over the symbols, so a seeded run reproduces the reference's candles."""

    def __init__(self, rng=None, batch_id: Optional[int] = None, whole_frames: bool = True):
        self.batch_id = int(time.time()) if batch_id is None else batch_id
        self.rng = rng
        # a stage hands its WHOLE result frame to the next one (hostpath.FrameSlices) instead of per-symbol frames that are
        # cut, column-selected and concatenated again: same frames out, the pandas work per symbol drops from ~2.4 ms to three
        # ``iloc`` slices.  ``False`` keeps the per-symbol hand-over (the two are held equal by the tests).
        self.whole_frames = whole_frames

    @staticmethod
    def _tiles(whole: Optional[dict], names: List[str], expect: List[str]):
        """``FrameSlices`` over a stage's whole result frame when the symbols that got a result are exactly ``expect``, in
        order, and their row ranges tile the frame; ``None`` otherwise (-> per-symbol hand-over)."""
        from .hostpath import FrameSlices
        if not whole or "frame" not in whole or len(expect) < 2:
            return None
        off = whole["off"]
        lens = np.diff(off)
        ok = whole.get("ok")
        got = (lens > 0) if ok is None else (lens > 0) & np.asarray(ok, dtype=bool)
        if len(names) != len(lens) or [n for n, g in zip(names, got) if g] != expect or int(lens[got].sum()) != len(whole["frame"]):
            return None
        return FrameSlices(whole["frame"], np.concatenate([[0], np.cumsum(lens[got])]))

    def run(self, frames: Mapping[str, pd.DataFrame]):
        from .data_bridge.ohlcv_converter import InterpolatedToOHLCVConverter

        t0 = time.time()
        # ---- Task 1 (complete_pipeline.py:49 default-constructs the interpolator) --------------------------------------
        w1 = {} if self.whole_frames else None
        interp, st1, sum1 = InterpolationBatch(batch_id=self.batch_id).run(frames, whole_out=w1)
        # ---- bridge: SELECT symbol, date, underlying_price, mark_price, index_price, volume ... ORDER BY date (:441-446) ----
        syms = list(frames)
        st2: List[dict] = []
        have = [s for s in syms if s in interp]
        conv = InterpolatedToOHLCVConverter(None, None, rng=self.rng)
        w2 = {} if self.whole_frames else None
        try:
            fs1 = self._tiles(w1, (w1 or {}).get("names", []), have)     # the bridge reads its six columns by name
            gen = conv._generate_many(fs1 if fs1 is not None else
                                      [interp[s][["symbol", "date", "underlying_price", "mark_price", "index_price", "volume"]] for s in have],
                                      "pipeline", whole_out=w2) if have else []
            err = None
        except Exception as e:                                                        # :470-471
            gen, err = [], str(e)
        minute: Dict[str, pd.DataFrame] = {}
        have_pos = {s: i for i, s in enumerate(have)}
        for s in syms:
            if s not in interp:
                st2.append({"symbol": s, "status": "skipped", "reason": "No interpolated data"})              # :451-452
            elif err is not None:
                st2.append({"symbol": s, "status": "error", "error": err})
            else:
                f = gen[have_pos[s]]["frame"]
                if f is None or f.empty:
                    st2.append({"symbol": s, "status": "error", "error": "OHLCV generation failed"})          # :456-457
                else:
                    minute[s] = f
                    st2.append({"symbol": s, "status": "success", "input_rows": len(interp[s]), "output_rows": len(f), "error": None})
        # ---- Task 2 (:641-684): N = 5 fixed, rows in the stored order, no sort, the caller's len < 5 early-out ----------
        st3: List[dict] = []
        todo = [s for s in syms if s in minute and len(minute[s]) >= 5]
        rec = CandleReconstructor("5min")
        w3 = {} if self.whole_frames else None
        try:
            fs2 = self._tiles(w2, have, todo) if err is None else None
            res = rec.reconstruct_batch(fs2 if fs2 is not None else
                                        [minute[s][["symbol", "timestamp", "open", "high", "low", "close", "volume"]] for s in todo],
                                        presorted=True, check_len=False, whole_out=w3, slices=False) if todo else []
            err = None
        except Exception as e:
            res, err = [], str(e)
        if w3 and "frame" in w3:                                                       # one frame for all: :693-704, :669-671 once
            bigc = w3["frame"][["timestamp", "open", "high", "low", "close", "volume", "symbol"]].copy()
            bigc["frequency"] = "5min"
            bigc["source_candles"] = 5
            bigc["batch_id"] = self.batch_id
            c_off, c_ok = w3["off"], w3["ok"]
            res = [None] * len(todo)
            for k, i in enumerate(w3["slots"]):                                        # slots: the positions in `todo` the offsets stand for
                if c_ok[k]:
                    res[i] = bigc.iloc[int(c_off[k]):int(c_off[k + 1])]
            prepared = True
        else:
            prepared = False
        candles: Dict[str, pd.DataFrame] = {}
        todo_pos = {s: i for i, s in enumerate(todo)}
        for s in syms:
            if s not in minute:
                st3.append({"symbol": s, "status": "skipped", "reason": "No minute candles found"})           # :657-658
            elif len(minute[s]) < 5:
                st3.append({"symbol": s, "status": "skipped", "reason": "Insufficient data for 5-min candles"})   # :660-661
            elif err is not None:
                st3.append({"symbol": s, "status": "error", "error": err})                                     # :683-684
            else:
                r = res[todo_pos[s]]
                if r is None or r.empty:
                    st3.append({"symbol": s, "status": "error", "error": "Reconstruction failed"})            # :665-666
                else:
                    if prepared:
                        out = r
                    else:
                        out = r[["timestamp", "open", "high", "low", "close", "volume", "symbol"]].copy()      # :693-704
                        out["frequency"] = "5min"                                                              # :669-671
                        out["source_candles"] = 5
                        out["batch_id"] = self.batch_id
                    candles[s] = out
                    st3.append({"symbol": s, "status": "success", "input_candles": len(minute[s]), "output_candles": len(out), "error": None})
        stages = {"interpolation": (st1, sum1), "bridge": (st2, _summary(self.batch_id, st2, t0)),
                  "reconstruction": (st3, _summary(self.batch_id, st3, t0))}
        return candles, minute, interp, stages


class ProductionInterpolationBatch(InterpolationBatch):
    """``ProductionBatchProcessor._process_single_symbol`` for every symbol of a batch (reference
    src/interpolation/optimized_batch_processor.py:267-331): symbols with more than ``streaming_threshold`` (10 000) rows
    are read ``fetch_size`` (5 000) rows at a time in date order and EVERY CHUNK IS INTERPOLATED ON ITS OWN (:333-360) — no
    interpolation across chunk boundaries, the guards of ``interpolate_symbol`` apply per chunk.  Small symbols and the
    chunks of large ones all go through one launch.

    The reference then evaluates ``len(interpolated_data)`` on the integer its streaming helper returned (:319), which
    raises, so after saving every chunk it reports the symbol as ``'error'``.  Here the symbol is reported as a success
    with the summed row count; ``mirror_reference_status=True`` reproduces the reference's status dictionary instead."""

    def __init__(self, streaming_threshold: int = 10_000, fetch_size: int = 5_000, mirror_reference_status: bool = False, **kw):
        super().__init__(**kw)
        self.streaming_threshold = streaming_threshold
        self.fetch_size = fetch_size
        self.mirror_reference_status = mirror_reference_status

    def run(self, frames: Mapping[str, pd.DataFrame]) -> Tuple[Dict[str, pd.DataFrame], List[dict], dict]:
        t0 = time.time()
        pieces: List[pd.DataFrame] = []
        owner: List[Tuple[str, bool]] = []
        for sym, f in frames.items():
            if f is None or f.empty:
                continue
            if len(f) > self.streaming_threshold:                                # :290-291
                g = f.sort_values("date", kind="stable") if "date" in f.columns else f   # the SQL's ORDER BY date
                for a in range(0, len(g), self.fetch_size):
                    pieces.append(g.iloc[a:a + self.fetch_size])
                    owner.append((sym, True))
            else:
                pieces.append(f)
                owner.append((sym, False))
        try:
            results = self.interpolator.interpolate_batch(pieces) if pieces else []
            err = None
        except Exception as e:
            results, err = [None] * len(pieces), str(e)
        if self.recompute_greeks and err is None:                              # one more launch over every piece's rows (f-4)
            try:
                self._add_greeks([r for r in results if r is not None])
            except Exception as e:
                results, err = [None] * len(pieces), str(e)
        elapsed = time.time() - t0
        chunks: Dict[str, List[pd.DataFrame]] = {}
        small: Dict[str, Optional[pd.DataFrame]] = {}
        save_err: Dict[str, str] = {}
        for (sym, streamed), res in zip(owner, results):
            if res is not None:
                res["batch_id"] = self.batch_id                                   # :312 / :352
                # the reference saves every chunk as it is produced (:353) and every small symbol (:312); `sink` stands there.
                # It ignores what _save_results_optimized returns; here a refused or failing save is reported on the symbol.
                if self.sink is not None and sym not in save_err:
                    try:
                        if not self.sink(sym, res):
                            save_err[sym] = "Failed to save to database"
                    except Exception as e:
                        save_err[sym] = str(e)
            if streamed:
                chunks.setdefault(sym, [])
                if res is not None:
                    chunks[sym].append(res)
            else:
                small[sym] = res
        outs: Dict[str, pd.DataFrame] = {}
        statuses: List[dict] = []
        for sym, f in frames.items():
            n_in = 0 if f is None else len(f)
            if n_in == 0:
                statuses.append({"symbol": sym, "status": "skipped", "reason": "No data found"})                 # :286-287
            elif err is not None:
                statuses.append({"symbol": sym, "status": "error", "error": err, "processing_time": elapsed})      # :324-331
            elif sym in save_err:
                statuses.append({"symbol": sym, "status": "error", "error": save_err[sym], "processing_time": elapsed})
            elif sym in chunks:
                parts = chunks[sym]
                if parts:
                    outs[sym] = pd.concat(parts, ignore_index=True)
                total = sum(len(p) for p in parts)
                if self.mirror_reference_status:
                    statuses.append({"symbol": sym, "status": "error", "error": "object of type 'int' has no len()",
                                     "processing_time": elapsed})
                else:
                    statuses.append({"symbol": sym, "status": "success", "input_rows": n_in, "output_rows": total,
                                     "processing_time": elapsed / max(1, len(frames))})
            elif small.get(sym) is None:
                statuses.append({"symbol": sym, "status": "skipped", "reason": "Interpolation failed"})           # :307-308
            else:
                outs[sym] = small[sym]
                statuses.append({"symbol": sym, "status": "success", "input_rows": n_in, "output_rows": len(small[sym]),
                                 "processing_time": elapsed / max(1, len(frames))})
        return outs, statuses, _summary(self.batch_id, statuses, t0)
