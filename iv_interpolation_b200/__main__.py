"""DB-free command line over the hot path: Parquet in, Parquet out.

    python -m iv_interpolation_b200 --task interpolation --in hourly.parquet --out minute_rows.parquet
    python -m iv_interpolation_b200 --task bridge        --in minute_rows.parquet --out minute_candles.parquet
    python -m iv_interpolation_b200 --task candles       --in minute_candles.parquet --out candles_5min.parquet
    python -m iv_interpolation_b200 --task pipeline      --in hourly.parquet --out candles_5min.parquet [--keep-stages]

The task names and the per-task summaries follow reference main.py:459-604 (``--task interpolation|bridge|candles|both|
pipeline``, ``--test`` = three symbols, main.py:267-343).  The reference reads and writes PostgreSQL and its ``candles``,
``both`` and ``pipeline`` tasks die on an undefined ``run_task_2`` (main.py:553-587); here every
task is defined, the tables are Parquet files with the reference's column names at the DECIMAL precisions of its DDL
(columnar.py), and all symbols of a file go through one launch per stage (tables.py).  ``both`` = interpolation then
bridge (two files out: --out and --out with ``.minute`` before the extension); ``pipeline`` = all three stages chained
on the device.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import pyarrow as pa
import pyarrow.compute as pc
import pyarrow.parquet as pq

from . import columnar, tables


def _read(path: str, test: bool) -> pa.Table:
    t = pq.read_table(path)
    if test:                                             # main.py:267-343 / complete_pipeline.py:759-763: three symbols
        syms = pc.unique(t["symbol"]).slice(0, 3)
        t = t.filter(pc.is_in(t["symbol"], value_set=syms))
    return t


def _n_symbols(t: pa.Table) -> int:
    return len(pc.unique(t["symbol"]))


def _summary(title: str, statuses: dict, verbose: bool) -> bool:
    ok = sum(s["status"] == "success" for s in statuses.values())
    err = sum(s["status"] == "error" for s in statuses.values())
    skip = sum(s["status"] == "skipped" for s in statuses.values())
    if verbose:
        for i, (sym, s) in enumerate(statuses.items(), 1):
            if s["status"] == "success":
                a = s.get("input_rows", s.get("input_candles")); b = s.get("output_rows", s.get("output_candles"))
                print(f"Processing {i}/{len(statuses)}: {sym}\n   SUCCESS: {a} → {b}")
            else:
                print(f"Processing {i}/{len(statuses)}: {sym}\n   FAILED: {s.get('error', s.get('reason', 'Unknown error'))}")
    print(f"\n{title} COMPLETE:\n   Success: {ok}\n   Errors: {err}\n   Skipped: {skip}")
    return ok > 0


def run_task_1(args) -> bool:
    print("TASK 1: IV Interpolation (Hourly → 1-minute)")
    t = _read(args.inp, args.test)
    print(f"PROCESSING: Found {_n_symbols(t)} symbols")
    out, st = tables.interpolate_table(t, method=args.method, min_points=args.min_points)
    out = out.append_column("batch_id", pa.array([args.batch_id] * len(out), pa.int64()))
    columnar.write_parquet(columnar.table_to_db_precision(out, columnar.INTERPOLATED_DECIMALS), args.out)
    ok = _summary("TASK 1", st, args.verbose)
    if ok:
        print(f"SUCCESS: Task 1 completed - Batch ID: {args.batch_id} - {len(out)} rows -> {args.out}")
    return ok


def run_data_bridge(args, inp=None, out_path=None) -> bool:
    print("BRIDGE: Converting interpolated data to OHLCV format")
    t = _read(inp or args.inp, args.test and inp is None)
    n_sym = _n_symbols(t)
    if n_sym == 0:
        print("FAILED: No interpolated data found for conversion\nHINT: Run Task 1 first: python -m iv_interpolation_b200 --task interpolation")
        return False
    print(f"CONVERTING: Found {n_sym} symbols with interpolated data")
    out = tables.bridge_table(t, strategy=args.strategy, seed=args.seed)
    columnar.write_parquet(columnar.table_to_db_precision(out, columnar.CANDLE_DECIMALS), out_path or args.out)
    done = _n_symbols(out)
    print(f"\nBRIDGE COMPLETE:\n   Success: {done}\n   Errors: 0\n   Skipped: {n_sym - done}")
    print("✅ Data bridge successful - ready for Task 2" if done else "❌ No symbols converted successfully")
    return done > 0


def run_task_2(args) -> bool:
    """The function reference main.py:553-587 calls and never defines."""
    print(f"TASK 2: Candle Reconstruction (1-minute → {args.frequency})")
    t = _read(args.inp, args.test)
    n_sym = _n_symbols(t)
    if n_sym == 0:
        print("FAILED: No symbols with minute candle data found")
        return False
    print(f"PROCESSING: Found {n_sym} symbols with minute data")
    out, st = tables.reconstruct_table(t, args.frequency)
    out = out.append_column("batch_id", pa.array([args.batch_id] * len(out), pa.int64()))
    columnar.write_parquet(columnar.table_to_db_precision(out, columnar.CANDLE_DECIMALS), args.out)
    return _summary("TASK 2", st, args.verbose)


def run_pipeline(args) -> bool:
    print("PIPELINE: Task 1 → Data Bridge → Task 2")
    t = _read(args.inp, args.test)
    want = ("interpolated", "minute", "candles") if args.keep_stages else ("candles",)
    res = tables.pipeline_table(t, min_points=args.min_points, strategy=args.strategy, target_frequency=args.frequency, seed=args.seed, want=want)
    c = res["counts"]
    stem, ext = os.path.splitext(args.out)
    if args.keep_stages:
        columnar.write_parquet(columnar.table_to_db_precision(res["interpolated"], columnar.INTERPOLATED_DECIMALS), f"{stem}.interpolated{ext}")
        columnar.write_parquet(columnar.table_to_db_precision(res["minute"], columnar.CANDLE_DECIMALS), f"{stem}.minute{ext}")
    columnar.write_parquet(columnar.table_to_db_precision(res["candles"], columnar.CANDLE_DECIMALS), args.out)
    n_ok = int((c["candles"] > 0).sum())
    print(f"\nPIPELINE COMPLETE:\n   Symbols: {len(c['status'])}\n   Success: {n_ok}\n   Skipped: {len(c['status']) - n_ok}\n"
          f"   Rows: {int(c['source_rows'].sum())} source → {int(c['interpolated'].sum())} interpolated → {int(c['minute'].sum())} minute candles → "
          f"{int(c['candles'].sum())} {args.frequency} candles")
    return n_ok > 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m iv_interpolation_b200", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--task", choices=("interpolation", "bridge", "candles", "both", "pipeline"), required=True)
    ap.add_argument("--in", dest="inp", required=True, help="input Parquet file (all symbols, any row order)")
    ap.add_argument("--out", required=True, help="output Parquet file")
    ap.add_argument("--test", action="store_true", help="first three symbols only (main.py --test)")
    ap.add_argument("--frequency", default="5min", help="target candle frequency (config.candle_reconstruction.target_frequency)")
    ap.add_argument("--method", default="linear", help="config.interpolation.method")
    ap.add_argument("--min-points", type=int, default=10, help="config.interpolation.min_data_points")
    ap.add_argument("--strategy", default="spread_simulation", help="config.data_bridge.conversion_strategy")
    ap.add_argument("--seed", type=int, default=0, help="seed of the bridge's device random tapes")
    ap.add_argument("--batch-id", type=int, default=int(time.time()), help="monitoring/progress.py:18-20")
    ap.add_argument("--keep-stages", action="store_true", help="pipeline: also write the interpolated rows and the minute candles")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args(argv)
    t0 = time.time()
    try:
        if args.task == "interpolation":
            ok = run_task_1(args)
        elif args.task == "bridge":
            ok = run_data_bridge(args)
        elif args.task == "candles":
            ok = run_task_2(args)
        elif args.task == "both":                        # main.py:560-566: Task 1, then the bridge on its output
            stem, ext = os.path.splitext(args.out)
            ok = run_task_1(args) and run_data_bridge(args, inp=args.out, out_path=f"{stem}.minute{ext}")
        else:
            ok = run_pipeline(args)
    except Exception as e:                                # main.py:600-602
        print(f"FAILED: {e}")
        return 1
    print(f"Total time: {time.time() - t0:.2f} s")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
