#!/usr/bin/env python
"""bench.py — the hot path of IV_INTERPOLATION on B200, measured.

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference [--steps K] [--warmup W]      (CPU arm: the oracle port on the host cores)

A step is one pass of the north_star pipeline over one SPX-like option chain (BASELINE.json configs[1]:
1 day of 1-s quotes, 2000 strikes x 30 expiries = 1.404 G quotes, 57.6 GB of inputs resident in HBM):

    Black-Scholes implied vol for every quote -> natural cubic spline across strike (512-pt grid) per
    (timestamp, expiry) -> across maturity (64-pt grid) -> 1-minute OHLC candles of every surface point.

`value` = quotes solved per second through the WHOLE step (IV solves/s); surface points/s rides along.
With N GPUs every rank owns one chain (one underlying x trading day): weak scaling, no collective in the
timed region.  `e2e` streams the same day from pinned host memory, chunk by chunk, and brings the candles
back.  `roofline` is kernel C's (the dominant kernel), from CUDA events around it in the timed steps.
`kernels` adds the same figures for every stage and for the two reference-pinned kernels (A: time-linear
interpolation + ffill; B: OHLCV bucket reduce) on a config-3-shaped tape.  One JSON line on stdout.

`--workload contracts` (and the `contracts` object of the default line) is the path the reference really has, on
BASELINE.json configs[2]: 100 underlyings x ~1200 contracts x 481 hourly rows (20 days), dealt to the N ranks by
underlying (LPT), every rank Task 1 -> data bridge -> Task 2 on its shard, the 5-minute candle tables gathered on
rank 0 over NCCL — ONE fixed tape whatever N (strong scaling).  Its CPU arm runs the reference's OWN classes
(oracle/_ref, copied untouched by oracle/make_ref.py) on one core and on a ProcessPoolExecutor of all cores, the
way reference src/interpolation/batch_processor.py:234-239 fans contracts out.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_emit = None
METRIC = "iv_solves_per_sec_full_path"
UNIT = "quotes/s"
N_EXP, N_STRIKE, N_T_DAY = 30, 2000, 23_400
N_GRID_K, N_GRID_M = 512, 64
BYTES_PER_QUOTE = 49                      # 5 x f64 + u8 read, f64 written
CHUNK_T = 120                             # e2e: stamps per host chunk (two 1-minute buckets)
# contracts workload (BASELINE.json configs[2])
CT_METRIC = "interpolated_points_per_sec_full_path"
CT_UNIT = "minute-rows/s"
CT_UNDERLYINGS, CT_ROWS, CT_SEED = 100, 481, 1000
CT_OUT_ROWS = (CT_ROWS - 1) * 60 + 1       # 28 801 minute rows per contract


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port, all host cores, bounded sample of the same workload
# ------------------------------------------------------------------------------------------------------
_CPU_SAMPLE = {}


def _cpu_worker(bounds):
    from oracle import surface_oracle as SO
    a, b = bounds
    c = _CPU_SAMPLE
    per_t = N_EXP * N_STRIKE
    sl = slice(a * per_t, b * per_t)
    res = SO.surface_pipeline(c["ts"][a:b], c["price"][sl], c["S"][sl], c["K"][sl], c["T"][sl], c["r"][sl],
                              c["is_call"][sl], N_EXP, N_STRIKE, c["grid_k"], c["grid_m"], 60_000_000, 1)
    return float(res["close"][~(res["close"] != res["close"])].sum())


def cpu_arm(steps: int, warmup: int, wall_target_s: float = 15.0):
    """Times oracle.surface_oracle.surface_pipeline (numpy/scipy port; the reference has no code for this
    path) over a pool of all host cores, the way the reference fans contracts out to a
    ProcessPoolExecutor (reference src/interpolation/batch_processor.py:234-239)."""
    import multiprocessing as mp

    import numpy as np

    from iv_interpolation_b200 import tapes
    from oracle import np_oracle as O

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_stamp_s = 0.55                                        # one core, measured in the build container
    n_t = int(min(480, max(cores, round(wall_target_s * cores / per_stamp_s / max(1, steps + warmup)))))
    n_t = max(cores, n_t // cores * cores)
    chain = tapes.make_quote_chain_np(7, n_t, N_EXP, N_STRIKE, O.bs_price, N_GRID_K, N_GRID_M)
    _CPU_SAMPLE.clear()
    _CPU_SAMPLE.update(chain)
    per = n_t // cores
    bounds = [(i * per, (i + 1) * per) for i in range(cores)]
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            chk = sum(pool.map(_cpu_worker, bounds))
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    n_quotes = n_t * N_EXP * N_STRIKE
    ms = 1e3 * sum(times) / len(times)
    return dict(value=n_quotes / (ms / 1e3), unit=UNIT, cores=cores, kind="port",
                sample=f"{n_t} of the day's {N_T_DAY} stamps x {N_EXP} expiries x {N_STRIKE} strikes = {n_quotes} quotes per step, "
                       f"{cores} worker processes (oracle/surface_oracle.py: vectorised safeguarded Newton + scipy CubicSpline "
                       f"+ numpy candles), checksum {chk:.6g}"), ms, n_quotes


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):   # one thread per worker process
        os.environ.setdefault(k, "1")
    if args.workload == "dropin":
        _emit({"impl": "reference", "dropin": cpu_dropin_ref()})
        return
    if args.workload == "contracts":
        cb = cpu_contracts_ref()
        if "unavailable" in cb:
            _emit({"impl": "reference", "unavailable": cb["unavailable"]})
            return
        n_c, _ = _ct_plan(args.ct_underlyings)
        _emit({"impl": "reference", "metric": CT_METRIC, "value": cb["value"], "unit": CT_UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 1,
               "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": _ct_workload_name(args.ct_underlyings, int(n_c.sum())),
                          "note": "the step is a bounded sample of the workload's contracts, see cpu_baseline.sample"},
               "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": CT_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0})
        return
    cb, ms, n_quotes = cpu_arm(args.steps, args.warmup)
    contracts_ref = None
    if not args.skip_contracts:
        try:
            contracts_ref = cpu_contracts_ref()
        except Exception as e:      # a reported baseline: never lose the line
            contracts_ref = {"failed": repr(e)}
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": _workload_name(N_T_DAY), "note": "each step is a bounded sample of the workload, see cpu_baseline.sample"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "contracts": {"metric": CT_METRIC, "unit": CT_UNIT, "cpu_baseline": contracts_ref,
                          "note": "the reference's OWN classes on the contracts workload's tape (the surface line above is the oracle port: "
                                  "the reference has no IV solver and no spline)"}}
    _emit(line)


def _workload_name(n_t):
    return (f"SPX-like chain (BASELINE.json configs[1]): {n_t} 1-s stamps x {N_EXP} expiries x {N_STRIKE} strikes "
            f"-> {N_GRID_M}x{N_GRID_K} surface per stamp -> 1-min OHLC candles")


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t_begin: float, t_end: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [ln for (t, ln) in self.lines if t_begin <= t <= t_end] or [ln for (_, ln) in self.lines[-3:]]
        sm, mx, reasons, pw = [], [], set(), []
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
class _Ctx:
    """One process per GPU: device, rank plumbing, barrier and max/sum over ranks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (there is no CPU path; use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL's INFO log (rank/communicator lines the driver counts) goes to stderr with everything else: stdout is
            # redirected in main(), the one JSON line is written through the saved descriptor
            os.environ["NCCL_DEBUG"] = os.environ.get("IVX_NCCL_DEBUG", "INFO")          # whatever the image presets (WARN) is overridden
            os.environ["NCCL_DEBUG_SUBSYS"] = os.environ.get("IVX_NCCL_DEBUG_SUBSYS", "INIT")
            # the final gather is point to point (rank r -> rank 0): give send/recv more than NCCL's default two channels per peer
            os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "16")
            os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "32")
            dist.init_process_group("nccl", device_id=self.dev)
        if args.gpus != self.world:
            if self.rank == 0:
                print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={self.world}; launch with torch.distributed.run", file=sys.stderr)
            args.gpus = self.world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x: float, op) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum_over_ranks(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def gather_floats(self, x: float):
        if self.world == 1:
            return [x]
        t = self.torch.zeros(self.world, dtype=self.torch.float64, device=self.dev)
        t[self.rank] = x
        self.dist.all_reduce(t)
        return t.tolist()

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def run_gpu(args):
    ctx = _Ctx(args)
    torch, dist = ctx.torch, ctx.dist
    world, rank, local, dev = ctx.world, ctx.rank, ctx.local, ctx.dev
    barrier, max_over_ranks = ctx.barrier, ctx.max_over_ranks

    from iv_interpolation_b200 import device as D, surface as SF, tapes

    if args.workload == "contracts":
        return run_gpu_contracts(ctx, args)

    # ---- the contracts workload first (its working set is released before the chain is built) ------------------------
    contracts = None
    if not args.skip_contracts:
        contracts = bench_contracts(ctx, args, _peaks()[0])
        if rank == 0 and world == 1 and not args.skip_cpu:
            contracts["cpu_baseline"] = _cpu_leg_subprocess("contracts", "cpu_baseline")
        torch.cuda.empty_cache()

    dropin = None
    if rank == 0 and world == 1 and not args.skip_dropin:
        try:
            dropin = bench_dropin(args)
            if not args.skip_cpu:
                dropin["cpu_baseline"] = _cpu_leg_subprocess("dropin", "dropin")
        except Exception as e:       # host-side figures: never lose the line
            dropin = {"failed": repr(e)}
        torch.cuda.empty_cache()

    # ---- size the chain to the GPU (the full day needs ~80 GB; shrink the day only if it does not fit) ----
    free, total = torch.cuda.mem_get_info(dev)
    per_stamp = N_EXP * N_STRIKE * (BYTES_PER_QUOTE + 8) + N_EXP * N_GRID_K * 8 + N_GRID_M * N_GRID_K * 8 + 1200   # + sigma_true
    n_t = args.n_t or N_T_DAY
    fit = int((free - (24 << 30)) // per_stamp) // 60 * 60       # leave room for the A/B sub-benchmarks and the e2e slots
    if n_t > fit:
        n_t = max(60, fit)
    spec = SF.ChainSpec(n_t, N_EXP, N_STRIKE, N_GRID_K, N_GRID_M, bucket_ticks=60 * tapes.US, min_count=1)
    chain = tapes.make_quote_chain(100 + rank, spec, dev)
    del chain.sigma_true
    torch.cuda.empty_cache()
    pipe = SF.SurfacePipeline(spec, dev, max_buckets=n_t // 60 + 2)
    pipe.plan_buckets(chain.ts)
    n_quotes = spec.n_quotes

    # ---- value: inputs resident in HBM ---------------------------------------------------------------------
    for _ in range(args.warmup):
        pipe.run(chain)
    ev = [{s: [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)] for s in pipe.STAGES}
          for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    D.reset_launch_count()
    t_begin = time.perf_counter()
    e0.record()
    for k in range(args.steps):
        res = pipe.run(chain, events=ev[k])
    e1.record()
    barrier()
    t_end = time.perf_counter()
    launches = D.launch_count()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = world * n_quotes / (ms_step / 1e3)
    stage_ms = {s: sum(ev[k][s][0].elapsed_time(ev[k][s][1]) for k in range(args.steps)) / args.steps for s in pipe.STAGES}
    n_buckets = res.n_buckets
    checksum = float(torch.nan_to_num(res.close).sum().item())
    nan_iv = float(torch.isnan(res.iv[: min(n_quotes, 1 << 26)]).double().mean().item())

    peak, peak_src = _peaks()
    alg = {
        "implied_vol": n_quotes * BYTES_PER_QUOTE,
        # knots: y only (8 B) when the solver has already compared the strikes with the ladders (banded path, x_verified), else x and y
        "spline_strike": n_quotes * (8 if pipe.banded else 16) + n_t * N_EXP * N_GRID_K * 8,
        "spline_maturity": n_t * (N_EXP + N_GRID_M) * N_GRID_K * 8,
        "candles": n_t * N_GRID_M * N_GRID_K * 8 + n_buckets * N_GRID_M * N_GRID_K * 32,
    }
    units = {"implied_vol": (n_quotes, "quotes/s"), "spline_strike": (n_t * N_EXP * N_GRID_K, "points/s"),
             "spline_maturity": (n_t * N_GRID_M * N_GRID_K, "points/s"), "candles": (n_t * N_GRID_M * N_GRID_K, "rows/s")}
    kernels = {}
    for s in pipe.STAGES:
        gbs = alg[s] / (stage_ms[s] / 1e3) / 1e9
        kernels[s] = {"ms": round(stage_ms[s], 4), "rate": units[s][0] / (stage_ms[s] / 1e3), "rate_unit": units[s][1],
                      "algorithmic_bytes": alg[s], "gbs": round(gbs, 1), "frac_hbm": round(gbs / peak, 4),
                      "share_of_step": round(stage_ms[s] / ms_step, 4)}
    dom = max(pipe.STAGES, key=lambda s: stage_ms[s])
    roofline = {"kernel": {"implied_vol": "k_bs_implied_vol_ladder2" if pipe.banded else "k_bs_implied_vol",
                           "spline_strike": "k_band_apply_mma" if pipe.banded else "k_spline_fit_eval",
                           "spline_maturity": "k_spline_strided", "candles": "k_ohlc_time_major"}[dom],
                # the figures below are against the HBM roofline (algorithmic bytes / time / measured copy peak); what actually bounds
                # the dominant kernel is said in `bound` and `limiter`
                "bound": "hbm", "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac_hbm"],
                "traffic": None, "peak_source": peak_src,
                "note": "achieved = algorithmic bytes per launch / CUDA-event duration of that launch inside the timed steps"}

    # DRAM traffic and pipe utilisation of the dominant kernel, from the committed ncu capture (per unit, scaled to this launch)
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")) as f:
            tr = json.load(f).get(roofline["kernel"])
        if tr:
            n_units = {"quote": n_quotes, "knot": n_quotes, "grid point": n_t * N_GRID_M * N_GRID_K, "surface row": n_t * N_GRID_M * N_GRID_K}[tr["unit"]]
            roofline["traffic"] = tr["bytes_per_unit"] * n_units
            roofline["traffic_source"] = f"profiles/ncu_traffic_r2.json: {tr['bytes_per_unit']} DRAM B per {tr['unit']} (ncu --set full) x {n_units}"
    except (OSError, ValueError, KeyError):
        pass
    if dom == "implied_vol":
        # ncu at the full day (profiles/ncu_r2_fullday_summary.json): 369 warp instructions per 32 quotes, ~121 of them fp64; an fp64
        # instruction holds the issue port for two cycles on B200 (16 fp64 lanes per SM sub-partition).  The one-quote-per-thread
        # form runs AT that limit (406 + 133 = 539 of 560 cycles); the two-quote form has fewer instructions and stalls a little more
        roofline["bound"] = "issue/fp64"
        roofline["hbm_frac"] = roofline["frac"]
        roofline["limiter"] = ("instruction issue, with fp64 instructions costing two issue cycles each (ncu at the full day, "
                               "profiles/ncu_r2_fullday_summary.json: issue slots 73 % busy + the second cycle of the fp64 instructions 24 %; "
                               "fp64 pipe 48 %, DRAM 44 %): 369 warp instructions per 32 quotes, ~121 of them fp64; 49.0 B per quote moved for 49 algorithmic")
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_r2_summary.json")) as f:
                for e in json.load(f):
                    if e["kernel"].startswith(roofline["kernel"]):
                        roofline["fp64_pipe_frac"] = round(float(e["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"].split()[0]) / 100, 4)
                        roofline["issue_active_frac"] = round(float(e["smsp__issue_active.avg.pct_of_peak_sustained_active"].split()[0]) / 100, 4)
                        break
        except (OSError, ValueError, KeyError):
            pass

    # ---- the same step in the chain layout (S per stamp, K and call/put per contract, T per (stamp, expiry)) -----
    sc = SF.StructuredChain.from_quotes(chain)
    for _ in range(args.warmup):
        pipe.run(sc)
    evc = [{s: [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)] for s in pipe.STAGES}
           for _ in range(args.steps)]
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    c0.record()
    for k in range(args.steps):
        resc = pipe.run(sc, events=evc[k])
    c1.record()
    barrier()
    ms_chain = max_over_ranks(c0.elapsed_time(c1)) / args.steps
    chain_layout = {"value": world * n_quotes / (ms_chain / 1e3), "unit": UNIT, "ms_per_step": ms_chain,
                    "stage_ms": {s: round(sum(evc[k][s][0].elapsed_time(evc[k][s][1]) for k in range(args.steps)) / args.steps, 4)
                                 for s in pipe.STAGES},
                    "input_bytes_per_quote": 8,
                    "checksum_close": float(torch.nan_to_num(resc.close).sum().item()),
                    "note": "k_bs_implied_vol_chain + shared-ladder spline path; same quotes, same candles to within the IV tolerance"}
    del resc

    # ---- fp64 FMA peak, measured in the same run (kernel C is fp64-pipe bound) -------------
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    n_thr, iters = 148 * 2048, 4096
    for _ in range(2):
        D.probe_fp64_fma(n_thr, iters, sink)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(); D.probe_fp64_fma(n_thr, iters, sink); p1.record(); torch.cuda.synchronize()
    fp64_tflops = n_thr * iters * 8 * 2 / (p0.elapsed_time(p1) / 1e3) / 1e12
    roofline["fp64_fma_peak_tflops_measured"] = round(fp64_tflops, 2)

    # ---- e2e: the same day streamed from pinned host memory ---------------------------------------------
    e2e = e2e_q = e2e_t = probe = None
    if not args.skip_e2e:
        probe = _pcie_probe(ctx)
        e2e = _bench_e2e(args, torch, SF, tapes, dev, chain, n_t, world, barrier, max_over_ranks, "chain")
        e2e_t = _bench_e2e(args, torch, SF, tapes, dev, chain, n_t, world, barrier, max_over_ranks, "ticks")
        e2e_q = _bench_e2e(args, torch, SF, tapes, dev, chain, n_t, world, barrier, max_over_ranks, "quotes")
        for e in (e2e, e2e_t, e2e_q):
            e["pcie_probe_h2d_gbs_per_rank"] = probe["h2d_gbs_per_rank"]
            e["pcie_of_probe"] = round(e["pcie_gbs"] / (sum(probe["h2d_gbs_per_rank"]) / world), 3)     # this rank's H2D rate / what the bare probe gets

    # ---- reference-pinned kernels A and B on a config-3-shaped tape, every rank its own contracts --------
    del pipe, res, chain, sc
    torch.cuda.empty_cache()
    stress = None
    if rank == 0 and not args.skip_stress:
        try:
            stress = bench_stress(args, torch, D, SF, tapes, dev, peak)
        except Exception as e:       # a reported extra: never lose the line
            stress = {"failed": repr(e)}
        torch.cuda.empty_cache()
    if not args.skip_kernels:
        ab = _bench_ab(args, torch, D, tapes, dev, peak, rank, world, barrier, max_over_ranks)
        kernels.update(ab)

    if rank == 0 and world == 1 and not args.skip_cpu and not args.skip_kernels:
        try:
            for k, v in _cpu_port_ab(tapes).items():
                kernels[k]["cpu_port"] = v
        except Exception as e:  # reported baselines only
            kernels["interp_linear_ffill"]["cpu_port"] = {"failed": repr(e)}

    cb = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cb = _cpu_leg_subprocess("surface", "cpu_baseline")
        if cb.get("unit") is None:
            cb["unit"] = UNIT

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": _workload_name(n_t), "quotes_per_gpu": n_quotes, "chains": world,
                           "l2": "inputs (57.6 GB per step at the full day) exceed the 126 MB L2; no flush needed",
                           "sharding": "one chain (underlying x trading day) per GPU, no collective in the timed region",
                           "full_day": n_t == N_T_DAY},
                "surface_points_per_sec": world * n_t * N_GRID_M * N_GRID_K / (ms_step / 1e3),
                "candles_per_step": world * n_buckets * N_GRID_M * N_GRID_K,
                "roofline": roofline, "kernels": kernels, "chain_layout_resident": chain_layout, "cpu_baseline": cb, "e2e": e2e, "e2e_tick_layout": e2e_t, "e2e_per_quote_layout": e2e_q, "pcie_probe": probe, "gpu_launches": launches * world,
                "clocks": clocks, "checksum_close": checksum, "nan_iv_frac": nan_iv, "contracts": contracts, "e2e_dropin": dropin, "kernels_stress": stress}
        _emit(line)
    ctx.close()


def _cpu_leg_subprocess(workload: str, key: str):
    """The CPU leg in a fresh interpreter (no CUDA context to fork, BLAS pinned to one thread per worker)."""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    try:
        pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", "1",
                             "--warmup", "0", "--skip-contracts"], env=env, capture_output=True, text=True, timeout=900)
        return json.loads(pr.stdout.strip().splitlines()[-1])[key]
    except Exception as e:  # the CPU leg is a reported baseline; its failure must not lose the GPU numbers
        return {"value": None, "cores": None, "kind": None, "sample": f"failed: {e!r}"}


def run_gpu_contracts(ctx, args):
    """`--workload contracts`: the contracts workload as the line's top level."""
    torch = ctx.torch
    peak, peak_src = _peaks()
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()
        time.sleep(0.25)
    t_begin = time.perf_counter()
    c = bench_contracts(ctx, args, peak)
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end) if ctx.rank == 0 else None
    cb = None
    if ctx.rank == 0 and ctx.world == 1 and not args.skip_cpu:
        cb = _cpu_leg_subprocess("contracts", "cpu_baseline")
    if ctx.rank == 0:
        c["roofline"]["peak_source"] = peak_src
        line = {"metric": CT_METRIC, "value": c["value"], "unit": CT_UNIT, "n_gpus": ctx.world, "steps": c["steps"], "warmup": c["warmup"],
                "ms_per_step": c["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": c["workload"], "sharding": c["sharding"],
                           "l2": "every stage streams inputs far larger than the 126 MB L2 (>= 2.9 GB per batch); no flush needed"},
                "roofline": c["roofline"], "stages": c["stages"], "cpu_baseline": cb, "e2e": c["e2e"], "gpu_launches": c["gpu_launches"],
                "clocks": clocks, "counts": {k: c[k] for k in ("interp_rows", "minute_candles", "candles_5min", "source_rows", "batches_per_rank")}}
        _emit(line)
    ctx.close()


def bench_stress(args, torch, D, SF, tapes, dev, peak):
    """The same pipeline on tapes that are NOT the friendliest possible (one tenth of the day each, times scaled to the day):
      missing    2 % of the prices are NaN: the affected slices leave the banded spline for the exact scan (squeeze + general solve);
      shifted    1 % of the stamps quote on a ladder shifted by half a strike step: those slices fail the bit-for-bit ladder check;
      iid        SURVEY.md 8(d) config 2 as written: sigma ~ U(0.05, 1.5), T ~ U(1/365, 2) years, K ~ U(0.5, 1.5) S per QUOTE
                 (IV kernel only: iid quotes form no smile to fit).
    Reported per tape: stage times per day, the fraction of warps the IV fast path deferred to its retry kernel, the fraction of
    slices the banded spline sent to the exact solve."""
    n_t = max(60, (args.n_t or N_T_DAY) // 10 // 60 * 60)
    scale = N_T_DAY / n_t
    spec = SF.ChainSpec(n_t, N_EXP, N_STRIKE, N_GRID_K, N_GRID_M, bucket_ticks=60 * tapes.US, min_count=1)
    out = {"stamps": n_t, "note": f"{n_t} stamps per tape, times scaled x{scale:.1f} to the {N_T_DAY}-stamp day"}

    def run(chain, label):
        pipe = SF.SurfacePipeline(spec, dev, max_buckets=n_t // 60 + 2, banded="auto")
        pipe.plan_buckets(chain.ts)
        for _ in range(2):
            pipe.run(chain)
            torch.cuda.synchronize()
        ev = {st: [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)] for st in pipe.STAGES}
        res = pipe.run(chain, events=ev)
        torch.cuda.synchronize()
        n_warps = (spec.n_quotes + 31) // 32
        ws_iv = pipe._ivl_ws if getattr(pipe, "_ivl_ws", None) is not None else pipe._iv_ws
        deferred = int(ws_iv[:4].view(torch.int32).item()) if ws_iv is not None else None
        fb = int(pipe._fb[1].item()) if getattr(pipe, "_fb", None) is not None and not pipe.banded_switched_off else None
        ok = ~torch.isnan(chain.price)
        iv_nan = float(torch.isnan(res.iv[ok]).double().mean().item())
        out[label] = {"stage_ms_per_day": {st: round(ev[st][0].elapsed_time(ev[st][1]) * scale, 3) for st in pipe.STAGES},
                      "iv_warps_deferred_frac": None if deferred is None else round(deferred / n_warps, 6),
                      "spline_slices_exact_frac": None if fb is None else round(fb / (n_t * N_EXP), 6),
                      "strike_spline_path": "scan kernel (the pipeline left the banded operators after its first run: most slices need the exact solve)"
                                            if pipe.banded_switched_off else "banded operators + exact solve of the flagged slices",
                      "iv_nan_frac_of_quoted": round(iv_nan, 6)}
        del pipe, res

    g = torch.Generator(device=dev); g.manual_seed(99)
    chain = tapes.make_quote_chain(203, spec, dev, drop_frac=0.001)
    run(chain, "missing_0p1pct")          # two gaps per 2000-knot slice on average: filled on the tensor path unless they sit close together or near an end
    del chain
    chain = tapes.make_quote_chain(201, spec, dev, drop_frac=0.02)
    run(chain, "missing_2pct")
    del chain
    chain = tapes.make_quote_chain(202, spec, dev)
    step = float(chain.K[1] - chain.K[0])
    per_t = N_EXP * N_STRIKE
    moved = torch.nonzero(torch.rand(n_t, generator=g, device=dev) < 0.01).flatten()
    moved = moved[moved > 0]                                 # the first stamp defines the ladders
    for t in moved.tolist():
        sl = slice(t * per_t, (t + 1) * per_t)
        chain.K[sl] += 0.5 * step
        chain.price[sl] = D.bs_price(chain.S[sl], chain.K[sl], chain.T[sl], chain.r[sl], chain.sigma_true[sl], chain.is_call[sl])
    run(chain, "shifted_ladder_1pct")
    out["shifted_ladder_1pct"]["stamps_moved"] = int(moved.numel())
    # iid config-2 quotes: per-quote sigma, T, K — the IV kernel alone
    n = spec.n_quotes
    f64 = dict(dtype=torch.float64, device=dev)
    S = torch.full((n,), 4500.0, **f64)
    K = S * (0.5 + torch.rand(n, generator=g, **f64))
    T = 1.0 / 365.0 + torch.rand(n, generator=g, **f64) * (2.0 - 1.0 / 365.0)
    sig = 0.05 + 1.45 * torch.rand(n, generator=g, **f64)
    r = torch.full((n,), 0.05, **f64)
    call = (torch.rand(n, generator=g, device=dev) < 0.5).to(torch.uint8)
    price = D.bs_price(S, K, T, r, sig, call)
    ws = torch.empty(D.lib().ivx_bs_implied_vol_workspace_bytes(n), dtype=torch.uint8, device=dev)
    iv = torch.empty(n, **f64)
    for _ in range(2):
        D.bs_implied_vol(price, S, K, T, r, call, out=iv, workspace=ws)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); D.bs_implied_vol(price, S, K, T, r, call, out=iv, workspace=ws); e1.record(); torch.cuda.synchronize()
    okq = ~torch.isnan(iv)
    err = (iv[okq] - sig[okq]).abs()
    out["iid_cfg2"] = {"implied_vol_ms_per_day": round(e0.elapsed_time(e1) * scale, 3), "frac_hbm": round(n * BYTES_PER_QUOTE / (e0.elapsed_time(e1) / 1e3) / 1e9 / peak, 4),
                       "iv_warps_deferred_frac": round(int(ws[:4].view(torch.int32).item()) / ((n + 31) // 32), 6),
                       "solved_frac": round(float(okq.double().mean().item()), 6),
                       "round_trip_within_1e-8_frac": round(float((err <= 1e-8).double().mean().item()), 6)}
    return out


def _pcie_probe(ctx):
    """Bare pinned-memory copies, all ranks at once: what the box's host side gives every GPU when N ranks pull together —
    the denominator of the end-to-end figures (1 GiB per direction, best of 4; 2 GiB of pinned memory per rank)."""
    torch = ctx.torch
    n = 1 << 27                                                      # 1 GiB of float64
    host = torch.empty(n, dtype=torch.float64).pin_memory()
    host.fill_(1.0)
    devbuf = torch.empty(n, dtype=torch.float64, device=ctx.dev)
    res = {}
    for name, (dst, src) in {"h2d": (devbuf, host), "d2h": (host, devbuf)}.items():
        best = float("inf")
        for it in range(5):
            ctx.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); dst.copy_(src, non_blocking=True); b.record(); torch.cuda.synchronize()
            if it:
                best = min(best, a.elapsed_time(b))
        res[name + "_gbs_per_rank"] = [round(x, 2) for x in ctx.gather_floats(n * 8 / (best / 1e3) / 1e9)]
    res["how"] = "torch copy_ of 1 GiB between pinned host memory and the device, every rank at the same time (barrier before each copy), best of 4"
    del host, devbuf
    return res


def _bench_e2e(args, torch, SF, tapes, dev, chain, n_t, world, barrier, max_over_ranks, layout):
    """Through the host-facing entry (surface.HostSurfaceStream): every step copies the whole day's quotes
    H2D from pinned memory, chunk by chunk, and the candles D2H.  The host tape holds RING distinct chunks
    (pinned) that are re-sent cyclically until the day's stamps are covered — every byte crosses PCIe.
    layout "chain": S per stamp, K and call/put per contract (sent once), T per (stamp, expiry) — 8 B per quote;
    layout "ticks": the chain layout with prices as int32 ticks of 1e-4 (the reference's DECIMAL(15,4)) — 4 B per quote;
    layout "quotes": every column per quote — 41 B per quote."""
    n_chunks = max(1, n_t // CHUNK_T)
    # the host tape: distinct pinned chunks re-sent cyclically; at least ~2 GB of them so that the tape cannot sit in the host's
    # last-level cache (432 MiB did in round 1)
    chunk_bytes = CHUNK_T * N_EXP * N_STRIKE * {"chain": 8, "ticks": 4, "quotes": 41}[layout]
    ring = min(n_chunks, max(8, -(-(2 << 30) // chunk_bytes)))
    cspec = SF.ChainSpec(CHUNK_T, N_EXP, N_STRIKE, N_GRID_K, N_GRID_M, bucket_ticks=60 * tapes.US, min_count=1)
    per = cspec.n_quotes
    pin = lambda t: t.cpu().pin_memory()
    host = []
    sc = SF.StructuredChain.from_quotes(chain) if layout in ("chain", "ticks") else None
    for i in range(ring):
        sl, tl = slice(i * per, (i + 1) * per), slice(i * CHUNK_T, (i + 1) * CHUNK_T)
        if layout in ("chain", "ticks"):
            cls = SF.HostChainChunk if layout == "chain" else SF.HostTickChainChunk
            price = sc.price[sl] if layout == "chain" else SF.HostTickChainChunk.encode(sc.price[sl])
            host.append(cls(ts=pin(sc.ts[tl]), price=pin(price), S_t=pin(sc.S_t[tl]),
                            T_te=pin(sc.T_te[i * CHUNK_T * N_EXP:(i + 1) * CHUNK_T * N_EXP]), r_t=pin(sc.r_t[tl])))
        else:
            cols = {name: pin(getattr(chain, name)[sl]) for name in SF.QUOTE_F64 + ("is_call",)}
            host.append(SF.HostQuoteChunk(ts=pin(chain.ts[tl]), **cols))
    chunks = [host[i % ring] for i in range(n_chunks)]
    nb_chunk = CHUNK_T // 60
    stream = SF.HostSurfaceStream(cspec, dev, chain.grid_k, chain.grid_m, max_buckets_per_chunk=nb_chunk + 1,
                                  K_ek=sc.K_ek if sc is not None else None, is_call_ek=sc.is_call if sc is not None else None,
                                  price_ticks=layout == "ticks")
    nb_total = n_chunks * (nb_chunk + 1)
    out = [torch.empty(nb_total * cspec.n_series, dtype=torch.float64).pin_memory() for _ in range(4)]
    out_ts = torch.empty(nb_total, dtype=torch.int64).pin_memory()
    out_cnt = torch.empty(nb_total, dtype=torch.int32).pin_memory()
    counts = [nb_chunk] * n_chunks                 # 1-s stamps starting on a minute: the host knows its bucket layout
    for _ in range(min(2, args.warmup)):
        stream.run(chunks[: min(8, n_chunks)], out, out_ts, out_cnt, bucket_counts=counts)
    steps = max(1, min(args.steps, 3)) if layout != "quotes" else 1      # the per-quote layout is the secondary figure
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        rows = stream.run(chunks, out, out_ts, out_cnt, bucket_counts=counts)
    e1.record()
    barrier()
    wall = (time.perf_counter() - t0) / steps
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    nq = n_chunks * per
    return {"value": world * nq / (ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": stream.h2d_bytes, "d2h_bytes_per_step": stream.d2h_bytes,
            "ms_per_step": ms, "wall_ms_per_step": wall * 1e3, "steps": steps, "quotes_per_step_per_gpu": nq, "buckets": rows,
            "layout": layout, "h2d_bytes_per_quote": round(stream.h2d_bytes / nq, 2),
            "api": "iv_interpolation_b200.surface.HostSurfaceStream.run (pinned host columns in, pinned host candles out)",
            "host_tape": f"{ring} distinct pinned chunks of {CHUNK_T} stamps ({host[0].nbytes() >> 20} MiB each, {ring * host[0].nbytes() >> 20} MiB in all) re-sent cyclically, "
                         f"{n_chunks} chunks per step; per-quote IVs stay on the device (candles are the result)",
            "pcie_gbs": stream.h2d_bytes / (ms / 1e3) / 1e9}


def _bench_ab(args, torch, D, tapes, dev, peak, rank, world, barrier, max_over_ranks):
    """Kernel A (reference src/interpolation/core.py:26-81) and kernel B (src/candle_reconstruction/core.py:59-89)
    on a config-3-shaped tape: every rank owns `--ab-contracts` contracts with 481 hourly rows (20 days) ->
    28 801 minute rows each.  Timed per launch with CUDA events, outputs preallocated (steady state of a pipeline);
    the plan step (reads the source rows only, then one 8-byte read-back) is reported separately."""
    import numpy as np
    n_contracts = args.ab_contracts
    tp = tapes.make_contract_tape(3 + rank, n_contracts, rows_per_contract=481)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    off, date = t(tp.off), t(tp.date)
    icols = [t(tp.cols[c]) for c in tapes.INTERP_COLS]
    fcols = [t(tp.cols[c]) for c in tapes.FILL_NUM_COLS]
    kw = dict(step=tapes.MINUTE_US, max_span=30 * tapes.DAY_US)
    n_it = args.warmup + args.steps

    def timed(fn):
        ms = []
        for it in range(n_it):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a0.record(); r = fn(); a1.record(); torch.cuda.synchronize()
            if it >= args.warmup:
                ms.append(a0.elapsed_time(a1))
        return max_over_ranks(sum(ms) / len(ms)), r

    plan_ms, plan = timed(lambda: D.interp_plan(off, date, icols, fcols, [None], **kw))
    res = D.interp_exec(plan)
    exec_ms, res = timed(lambda: D.interp_exec(plan, out=res))
    rows = res.total
    # exec writes per row: date 8 + 3 interp x 8 + 6 fill x 8 + symbol source index 4 + is_interpolated 1 = 85 B;
    # the source side (80 B per source row, 60x fewer rows) is read by the plan step
    bytes_exec = rows * (8 + 24 + 48 + 4 + 1)
    out = {}

    def entry(ms, nbytes, nrows, note):
        gbs = nbytes / (ms / 1e3) / 1e9
        return {"ms": round(ms, 4), "rate": world * nrows / (ms / 1e3), "rate_unit": "minute-rows/s", "algorithmic_bytes": nbytes,
                "gbs": round(gbs, 1), "frac_hbm": round(gbs / peak, 4), "note": note}

    out["interp_linear_ffill"] = entry(exec_ms, bytes_exec, rows,
                                       f"k_interp_exec: {n_contracts} contracts x 481 hourly rows -> {rows} minute rows per GPU, 85 B written per row")
    out["interp_plan"] = entry(plan_ms, tp.n_rows * 80, rows,
                               f"k_interp_plan + 2 scans over {tp.n_rows} source rows (80 B each); rate is quoted in output rows planned")
    S, vol = res.interp[1], res.fill[4]
    # the step between A and B in the reference (data_bridge/ohlcv_converter.py:138-263, default strategy): minute rows ->
    # synthetic 1-min candles, uniform/exponential tapes resident on the device (5 uniforms per emitted row)
    bp_ms, bplan = timed(lambda: D.bridge_plan(res.out_off, S, vol))
    out["bridge_plan"] = entry(bp_ms, rows * 16, rows, "ivx_bridge_plan: per-segment counts, two scans, per-symbol offsets (16 B read per row) + a 16-byte read-back")
    u_tape = torch.rand(5 * bplan.total, dtype=torch.float64, device=dev)
    e_tape = torch.empty(max(1, bplan.total_e), dtype=torch.float64, device=dev).exponential_()
    br_ms, br = timed(lambda: D.bridge_ohlcv(res.out_off, S, vol, bplan, "spread_simulation", u=u_tape, e=e_tape))
    bytes_br = rows * 16 + br.total * (40 + 56) + bplan.total_e * 8
    out["bridge_ohlcv"] = entry(br_ms, bytes_br, rows,
                                f"k_bridge_ohlcv_seg + k_bridge_check (spread_simulation): {rows} minute rows -> {br.total} synthetic 1-min candles per GPU, "
                                "16 B read per row + 40 B of uniform tape and 56 B written per candle; output allocation inside the timed call")
    del u_tape, e_tape, br
    # five DISTINCT input columns (the same array passed four times would be served from L2)
    hi_, lo_, cl_ = S + 3.0, S - 3.0, S + 0.5
    b_ms, cb = timed(lambda: D.ohlcv_reduce(res.out_off, res.ts, S, hi_, lo_, cl_, vol, bucket_ticks=5 * tapes.MINUTE_US, min_count=5, trim=False))
    n_candles = int(cb._total_dev.item())
    bytes_b = rows * 48 + n_candles * 52
    out["ohlcv_reduce"] = entry(b_ms, bytes_b, rows,
                                f"{rows} minute rows -> {n_candles} 5-min candles per GPU: memsets + k_ohlcv_reduce + k_ohlcv_finalize, "
                                "48 B read per row + 52 B written per candle")
    return out


def _cpu_port_ab(tapes):
    """The oracle restatements of kernels A, B and the bridge (oracle/np_oracle.py, oracle/bridge_oracle.py — the reference's
    pandas path cannot travel to the GPU box) timed on ONE host core on a bounded sample of the same config-3-shaped tape:
    reported next to the GPU figures, as SURVEY.md 8(d) asks."""
    import time
    import numpy as np
    from oracle import bridge_oracle as BO, np_oracle as O
    n_c = 12
    tp = tapes.make_contract_tape(3, n_c, rows_per_contract=481)
    kw = dict(step=tapes.MINUTE_US, max_span=30 * tapes.DAY_US)
    outs, rows = [], 0
    t0 = time.perf_counter()
    for i in range(n_c):
        a, b = int(tp.off[i]), int(tp.off[i + 1])
        w = O.interp_linear_ffill(tp.date[a:b], [tp.cols[c][a:b] for c in tapes.INTERP_COLS], [tp.cols[c][a:b] for c in tapes.FILL_NUM_COLS],
                                  [np.ones(b - a, bool)], **kw)
        outs.append(w)
        rows += len(w["interp"][0])
    t_a = time.perf_counter() - t0
    t0 = time.perf_counter()
    for w in outs:
        S = w["interp"][1]
        O.ohlcv_reduce(w["date"], S, S + 3.0, S - 3.0, S + 0.5, w["fill"][4], bucket_ticks=5 * tapes.MINUTE_US, min_count=5)
    t_b = time.perf_counter() - t0
    rs = np.random.RandomState(1)
    t0 = time.perf_counter()
    br_rows = 0
    for w in outs[:2]:
        S, vol = w["interp"][1], w["fill"][4]
        u, e, z = BO.draw_tapes(rs, "spread_simulation", S, vol)
        BO.bridge_ohlcv("spread_simulation", S, vol, u, e, z)
        br_rows += len(S)
    t_br = time.perf_counter() - t0
    mk = lambda r, t, what: {"rate": r / t, "rate_unit": "minute-rows/s", "cores": 1, "kind": "port", "sample": what}
    return {"interp_linear_ffill": mk(rows, t_a, f"oracle/np_oracle.interp_linear_ffill on {n_c} contracts x 481 hourly rows -> {rows} minute rows"),
            "ohlcv_reduce": mk(rows, t_b, f"oracle/np_oracle.ohlcv_reduce on the same {rows} minute rows, 5-min buckets"),
            "bridge_ohlcv": mk(br_rows, t_br, f"oracle/bridge_oracle.bridge_ohlcv (row loop, like the reference's iterrows) on {br_rows} minute rows")}


# ------------------------------------------------------------------------------------------------------
# contracts workload (BASELINE.json configs[2]): the path the reference has — Task 1 -> bridge -> Task 2
# ------------------------------------------------------------------------------------------------------
def _ct_plan(n_und: int):
    """Contracts per underlying of THE fixed tape (seeded, the same whatever N) and the cost of each underlying."""
    import numpy as np
    n_c = np.random.default_rng(CT_SEED).integers(900, 1501, size=n_und)          # 1200 on average
    return n_c, n_c.astype(np.float64) * CT_OUT_ROWS


def _ct_workload_name(n_und: int, n_contracts: int, args=None) -> str:
    if args is not None and _ct_is_replay(args):
        return (f"replay (BASELINE.json configs[4]): {n_und} underlyings x ~1200 contracts ({n_contracts} in all), {args.ct_windows} consecutive "
                f"20-day windows of hourly rows (~{args.ct_windows * 20} trading days), sparse: {args.ct_nan_frac:.0%} of the IVs NaN and "
                f"{args.ct_drop_frac:.0%} of the hourly rows missing -> minute rows -> synthetic 1-min candles -> 5-min candles")
    return (f"contracts (BASELINE.json configs[2]): {n_und} underlyings x ~1200 contracts ({n_contracts} in all) x {CT_ROWS} hourly rows "
            f"(20 days) -> {CT_OUT_ROWS} minute rows each -> synthetic 1-min candles -> 5-min candles")


def _ct_is_replay(args) -> bool:
    return args.ct_windows > 1 or args.ct_drop_frac > 0.0 or args.ct_nan_frac != 0.05


def _ct_build_shard(mine, n_c, pin=True, nan_frac=0.05, drop_frac=0.0):
    import numpy as np
    from iv_interpolation_b200 import contracts as CT, tapes
    first = np.concatenate([[0], np.cumsum(n_c)])
    parts = []
    for u in mine:
        tp = tapes.make_underlying_tape(CT_SEED + 1 + int(u), int(n_c[u]), CT_ROWS, underlying=f"u{u}", nan_frac=nan_frac, drop_frac=drop_frac)
        parts.append(CT.HostContractShard.from_tape(tp, ids=first[u] + np.arange(n_c[u]), pin=False))
    return CT.HostContractShard.concat(parts, pin=pin)


def bench_contracts(ctx, args, peak):
    """Strong scaling of one fixed tape: the underlyings are dealt to the ranks longest-first (sharding.lpt_partition),
    every rank runs Task 1 -> bridge -> Task 2 on its shard in batches, the 5-minute candle tables are gathered on rank 0
    over NCCL.  `resident`: the shard's source columns already in HBM.  `e2e`: pinned host columns in (H2D inside the timed
    region), every batch's candle table back to pinned host memory (D2H inside), then the gather."""
    import numpy as np
    torch, dist, dev, rank, world = ctx.torch, ctx.dist, ctx.dev, ctx.rank, ctx.world
    from iv_interpolation_b200 import contracts as CT, device as D, sharding as SH, tapes

    n_und = args.ct_underlyings
    n_c, costs = _ct_plan(n_und)
    shards = SH.lpt_partition(costs, world)
    mine = shards[rank]
    t0 = time.perf_counter()
    shard = _ct_build_shard(mine, n_c, nan_frac=args.ct_nan_frac, drop_frac=args.ct_drop_frac)
    t_tape = time.perf_counter() - t0
    n_win = max(1, args.ct_windows)            # replay: the same contracts over consecutive 20-day windows (timestamps shifted per window)
    pipe_fused = not args.ct_two_step_bridge
    stream = CT.HostContractStream(dev, max_rows_per_batch=args.ct_batch_rows, seed=7, fused_tapes=pipe_fused, want_row=not pipe_fused)
    cuts = stream.batches(shard)
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 2))

    # ---- resident: source columns of the whole shard on the device -------------------------------------------------
    d_off = shard.off.to(dev)
    d_date = shard.date.to(dev)
    d_cols = {c: shard.cols[c].to(dev) for c in CT.SRC_COLS}
    views = []
    win_ticks = (CT_ROWS - 1) * tapes.HOUR_US + tapes.HOUR_US          # one window = 20 days + the hour to the next window's first stamp
    d_dates = [d_date if wdx == 0 else d_date + wdx * win_ticks for wdx in range(n_win)]
    for wdx in range(n_win):
        for a, b in cuts:
            r0, r1 = int(shard.off[a]), int(shard.off[b])
            views.append(((d_off[a:b + 1] - r0).contiguous(), d_dates[wdx][r0:r1], [d_cols[c][r0:r1] for c in tapes.INTERP_COLS],
                          [d_cols[c][r0:r1] for c in tapes.FILL_NUM_COLS], int(shard.ids[a])))

    def resident(ev=None):
        rows = minute = candles = 0
        for k, (o, d, ic, fc, first) in enumerate(views):
            res = stream.pipe.run(o, d, ic, fc, first_contract=first, events=ev[k] if ev is not None else None)
            rows += res.interp.total; minute += res.minute.total; candles += res.candles.total
        return rows, minute, candles

    mk_ev = lambda: [{st: [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)] for st in CT.STAGES} for _ in views]
    for _ in range(warmup):
        resident()
    evs = [mk_ev() for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    D.reset_launch_count()
    e0.record()
    for k in range(steps):
        rows, minute, candles = resident(evs[k])
    e1.record()
    ctx.barrier()
    launches = D.launch_count()
    ms_res = ctx.max_over_ranks(e0.elapsed_time(e1)) / steps
    tot = lambda x: int(ctx.sum_over_ranks(float(x)))
    rows_all, minute_all, candles_all = tot(rows), tot(minute), tot(candles)
    stage_ms = {st: sum(ev[st][0].elapsed_time(ev[st][1]) for step in evs for ev in step) / steps for st in CT.STAGES}
    n_src = int(shard.off[-1]) * n_win
    m_e = 0.03 * minute                       # ~3 % of the volumes are NULL in the tape: exponential draws
    fused = pipe_fused
    alg = {"interp_plan": n_src * 80, "interp_exec": rows * 85, "bridge_plan": rows * 16,
           "bridge_tapes": 0 if fused else (5 * minute + m_e) * 8,
           # fused: base + volume + stamp read per row, six columns + stamp written per candle (no tape, no row index)
           "bridge_ohlcv": rows * 24 + minute * (48 + 8) if fused else rows * 16 + minute * (40 + 56 + 8) + m_e * 8,
           "candle_ts": 0 if fused else minute * 24,
           "ohlcv_reduce": minute * 48 + candles * 52}
    kern = {"interp_plan": "k_interp_plan + k_scan_i64 x2", "interp_exec": "k_interp_exec", "bridge_plan": "k_bridge_count_seg + scans + k_bridge_sym_off",
            "bridge_tapes": "none: the draws are made inside the bridge kernel (Philox)" if fused else "k_random_tape x2",
            "bridge_ohlcv": "k_bridge_ohlcv_seg<GEN> + k_bridge_check" if fused else "k_bridge_ohlcv_seg + k_bridge_check",
            "candle_ts": "none: the bridge kernel copies the stamps" if fused else "k_gather<i64>",
            "ohlcv_reduce": "k_ohlcv_reduce + k_ohlcv_finalize"}
    stages = {}
    for st in CT.STAGES:
        gbs = alg[st] / (stage_ms[st] / 1e3) / 1e9 if stage_ms[st] > 0 else 0.0
        stages[st] = {"ms": round(stage_ms[st], 4), "kernels": kern[st], "algorithmic_bytes": int(alg[st]), "gbs": round(gbs, 1),
                      "frac_hbm": round(gbs / peak, 4), "share_of_step": round(stage_ms[st] / (e0.elapsed_time(e1) / steps), 4)}
    dom = max(CT.STAGES, key=lambda st: stage_ms[st])
    del views, d_off, d_date, d_dates, d_cols
    torch.cuda.empty_cache()

    # ---- e2e: pinned host columns in, candle tables out to pinned host memory, then the gather over NCCL ------------
    def e2e_step(gev=None):
        # replay: every window is sent again (same bytes; the time shift does not change the work), gathered and dropped — a year of
        # candle tables does not have to fit the host
        out, n_got, g_ms = None, 0, 0.0
        for _w in range(n_win):
            out = stream.run(shard)
            tab = {k: torch.cat([t[k] for t in out["tables"]]) for k in out["tables"][0]} if out["tables"] else {}
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            got = SH.gather_tables(tab, dst=0) if world > 1 and tab else tab
            g1.record()
            n_got += int(got["timestamp"].numel()) if got is not None and "timestamp" in got else 0
            if gev is not None:
                g1.synchronize()
                g_ms += g0.elapsed_time(g1)
            if n_win > 1:
                out["tables"] = None
                del tab, got
        if gev is not None:
            gev.append(g_ms)
        return out, n_got

    for _ in range(warmup):
        e2e_step()
    gevs = [[] for _ in range(steps)]
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    tw = time.perf_counter()
    f0.record()
    for k in range(steps):
        out, n_got = e2e_step(gevs[k])
    f1.record()
    ctx.barrier()
    wall = (time.perf_counter() - tw) / steps
    ms_e2e = ctx.max_over_ranks(f0.elapsed_time(f1)) / steps
    ms_gather = ctx.max_over_ranks(sum(g[0] for g in gevs) / steps)
    h2d, d2h = tot(stream.h2d_bytes) * n_win, tot(stream.d2h_bytes) * n_win      # the stream counts one pass over the shard
    if rank == 0 and world > 1 and n_got != candles_all:
        raise SystemExit(f"bench.py: the gather delivered {n_got} candles, the ranks produced {candles_all}")
    per_rank_rows = ctx.gather_floats(float(rows))
    res = {
        "metric": CT_METRIC, "unit": CT_UNIT, "scaling": "strong",
        "workload": _ct_workload_name(n_und, int(n_c.sum()), args),
        "windows": n_win, "nan_iv_frac": args.ct_nan_frac, "missing_row_frac": args.ct_drop_frac,
        "value": rows_all / (ms_res / 1e3), "ms_per_step": ms_res, "steps": steps, "warmup": warmup,
        "interp_rows": rows_all, "minute_candles": minute_all, "candles_5min": candles_all,
        "source_rows": tot(n_src), "batches_per_rank": len(cuts),
        "sharding": {"unit": "underlying", "policy": "sharding.lpt_partition (longest first; reference analogue optimized_batch_processor.py:123-164)",
                     "imbalance_max_over_mean": round(SH.imbalance(costs, shards), 4), "underlyings_per_rank": [len(x) for x in shards],
                     "minute_rows_per_rank": [int(x) for x in per_rank_rows]},
        "stages": stages,
        "roofline": {"kernel": kern[dom], "stage": dom, "bound": "issue" if (dom == "bridge_ohlcv" and pipe_fused) else "hbm",
                     "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": stages[dom]["frac_hbm"], "traffic": None,
                     **({"hbm_frac": stages[dom]["frac_hbm"], "issue_active_frac": 0.72, "fp64_pipe_frac": 0.31,
                         "limiter": "instruction issue: with the random draws made inside the kernel 26 % of its instructions are Philox and the "
                                    "53-bit assembly (profiles/ncu_r2e_bridge_philox_summary.json, srcprof_r2e_bridge_philox.txt); `achieved`/`frac` stay "
                                    "the algorithmic bytes (24 B read per row, 56 B written per candle) over the stage time against the HBM copy peak"}
                        if (dom == "bridge_ohlcv" and pipe_fused) else {})},
        "e2e": {"value": rows_all / (ms_e2e / 1e3), "unit": CT_UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "wall_ms_per_step": wall * 1e3, "gather_ms": ms_gather,
                "gather": ("sharding.gather_tables over NCCL: the 5-min candle tables of all ranks land on rank 0 (point to point, "
                           f"{candles_all * 56 >> 20} MiB)") if world > 1 else "single rank: nothing to gather",
                "api": "iv_interpolation_b200.contracts.HostContractStream.run (pinned host source columns in, 5-min candle tables out to pinned host memory)",
                "result": "the 5-minute candles (8 columns); the interpolated rows and the synthetic 1-minute candles stay on the device"},
        "gpu_launches": launches * world, "tape_build_s": round(t_tape, 2),
        "data": "synthetic (tapes.make_underlying_tape, seeds fixed per underlying); bridge random draws from the device counter generator (Philox, " + ("inside the bridge kernel" if pipe_fused else "written as tapes") + ")",
    }
    del out, stream
    torch.cuda.empty_cache()
    return res


# ---- the reference's own classes on the host cores (oracle/_ref) ---------------------------------------------------
def _ref_contract_worker(job):
    """One contract through the reference's three classes, as reference batch_processor.py:67-142 -> data_bridge ->
    candle_reconstruction do it (minus the database): returns rows and seconds per stage."""
    seed, idx = job
    import numpy as np
    from iv_interpolation_b200 import tapes
    from oracle import reference_loader as RL
    tp = _REF_TAPE.get(seed)
    if tp is None:
        tp = _REF_TAPE[seed] = tapes.make_underlying_tape(seed, 64, CT_ROWS, underlying="u0")
    df = tp.frame(idx)
    t0 = time.perf_counter()
    a = RL.reference_interpolate(df)                                          # interpolation/core.py:16-85 (fillna shim)
    t1 = time.perf_counter()
    if a is None:
        return 0, 0, 0, t1 - t0, 0.0, 0.0
    b_in = a[["symbol", "date", "underlying_price", "mark_price", "index_price", "volume", "is_interpolated"]].rename(columns={"date": "timestamp"})
    t1 = time.perf_counter()
    m = RL.reference_converter()._generate_ohlcv_from_interpolated(b_in)       # data_bridge/ohlcv_converter.py:138-187
    t2 = time.perf_counter()
    c = RL.reference_reconstruct(m[["symbol", "timestamp", "open", "high", "low", "close", "volume"]])   # candle core.py:36-106
    t3 = time.perf_counter()
    return len(a), len(m), 0 if c is None else len(c), t1 - t0, t2 - t1, t3 - t2


_REF_TAPE = {}


def cpu_contracts_ref(budget_s: float = 24.0):
    """The reference's classes (oracle/_ref or /root/reference) on contracts of the same tape: one core, then a
    ProcessPoolExecutor(max_workers=os.cpu_count()) with one future per contract (batch_processor.py:234-239)."""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    from oracle import reference_loader as RL
    if not RL.available():
        return {"unavailable": "neither /root/reference nor oracle/_ref is present (run oracle/make_ref.py in the build container)"}
    RL.quiet()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    seed = CT_SEED + 1
    t0 = time.perf_counter()
    one = [_ref_contract_worker((seed, 0))]
    per_contract = time.perf_counter() - t0
    n1 = int(max(1, min(6, budget_s * 0.35 // per_contract - 1)))
    one += [_ref_contract_worker((seed, i)) for i in range(1, n1 + 1)]
    t_one = time.perf_counter() - t0
    rows_one = sum(r[0] for r in one)
    k = int(max(1, min(8, budget_s * 0.5 // per_contract)))
    jobs = [(seed, i % 64) for i in range(cores * k)]
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("fork")) as ex:
        list(ex.map(_ref_contract_worker, [(seed, 0)] * cores))               # start the workers, import pandas, build the tape
        t0 = time.perf_counter()
        res = list(ex.map(_ref_contract_worker, jobs))
        t_pool = time.perf_counter() - t0
    rows_pool = sum(r[0] for r in res)
    st = lambda rs, j: sum(r[j] for r in rs)
    return {
        "value": rows_pool / t_pool, "unit": CT_UNIT, "cores": cores, "kind": "reference",
        "source": f"{RL.where()}: IVInterpolator.interpolate_symbol -> InterpolatedToOHLCVConverter._generate_ohlcv_from_interpolated -> "
                  "CandleReconstructor.reconstruct_symbol_candles, the reference's own modules (oracle/make_ref.py copies them untouched), "
                  "Task 1 under the pandas-3 fillna(method=) shim (SURVEY.md 0.3); no database",
        "sample": f"{len(jobs)} contracts x {CT_ROWS} hourly rows ({rows_pool} minute rows) of the workload's tape, ProcessPoolExecutor({cores}), "
                  f"one future per contract; {t_pool:.1f} s",
        "one_core": {"value": rows_one / t_one, "unit": CT_UNIT, "contracts": len(one), "seconds": round(t_one, 2),
                     "stage_rows_per_s": {"interpolate_symbol": rows_one / max(1e-9, st(one, 3)), "bridge": st(one, 1) / max(1e-9, st(one, 4)),
                                          "reconstruct_symbol_candles": st(one, 1) / max(1e-9, st(one, 5))},
                     "stage_seconds": {"interpolate_symbol": round(st(one, 3), 3), "bridge": round(st(one, 4), 3),
                                       "reconstruct_symbol_candles": round(st(one, 5), 3)}},
        "pool_stage_seconds_summed": {"interpolate_symbol": round(st(res, 3), 2), "bridge": round(st(res, 4), 2),
                                      "reconstruct_symbol_candles": round(st(res, 5), 2)},
        "candles_5min_per_contract": res[0][2], "ms_per_step": t_pool * 1e3,
    }


# ------------------------------------------------------------------------------------------------------
# drop-in classes, frames in -> frames out (BASELINE.json configs[0]: 2 400 contracts x 25 hourly rows, one day)
# ------------------------------------------------------------------------------------------------------
DROPIN_CONTRACTS = 2400


def _dropin_frames(n):
    from iv_interpolation_b200 import tapes
    tp = tapes.make_contract_tape(0, n)
    return [tp.frame(i) for i in range(tp.n_sym)]


def bench_dropin(args):
    """What a user of the reference swaps in: IVInterpolator.interpolate_batch, InterpolatedToOHLCVConverter.convert_frames,
    CandleReconstructor.reconstruct_batch and PipelineBatch.run — DataFrames in, DataFrames out, wall clock, on the config-1
    batch.  The CPU comparator (cpu_dropin_ref) loops the reference's own classes over frames of the same tape."""
    import numpy as np
    import pandas as pd
    from iv_interpolation_b200.batch import PipelineBatch
    from iv_interpolation_b200.candle_reconstruction.core import CandleReconstructor
    from iv_interpolation_b200.data_bridge.ohlcv_converter import InterpolatedToOHLCVConverter
    from iv_interpolation_b200.interpolation.core import IVInterpolator
    import logging
    logging.getLogger("iv_interpolation_b200").setLevel(logging.ERROR)
    frames = _dropin_frames(args.dropin_contracts)
    out = {"workload": f"config 1 (BASELINE.json configs[0]): {len(frames)} contracts x 25 hourly rows -> 1 441 minute rows each; DataFrames in, DataFrames out",
           "timer": "time.perf_counter around the call (host work, H2D, kernels, D2H, frame assembly), best of 3 after one warm-up"}

    def best(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
        return min(ts), r

    itp = IVInterpolator()
    splits = []

    def timed_interp():
        splits.append({})
        return itp.interpolate_batch(frames, timing=splits[-1])
    t_a, res_a = best(timed_interp)
    rows = sum(len(r) for r in res_a if r is not None)
    timing = min(splits[1:], key=lambda d: sum(d.values()))             # the split of the fastest timed call (the warm-up call left out)
    reps_done = 1
    out["interpolate_batch"] = {"api": "IVInterpolator.interpolate_batch(frames)", "rows_out": rows, "seconds": round(t_a, 4), "rate": rows / t_a,
                                "rate_unit": "minute-rows/s", "frames_ok": sum(r is not None for r in res_a),
                                "split_s": {k: round(v / reps_done, 4) for k, v in timing.items()}}
    one = frames[7]
    t_1, _ = best(lambda: itp.interpolate_symbol(one), reps=5)
    out["interpolate_symbol"] = {"api": "IVInterpolator.interpolate_symbol(frame) — one contract per call, as the reference's drivers call it",
                                 "seconds_per_call": round(t_1, 5), "rate": 1441 / t_1, "rate_unit": "minute-rows/s"}
    ok = [r for r in res_a if r is not None]
    b_in = [r[["symbol", "date", "underlying_price", "mark_price", "index_price", "volume", "is_interpolated"]].rename(columns={"date": "timestamp"}) for r in ok]
    conv = InterpolatedToOHLCVConverter(None, None, rng=np.random.RandomState(1))
    t_br, res_br = best(lambda: conv.convert_frames(b_in), reps=2)
    minute = [r["ohlcv"][["symbol", "timestamp", "open", "high", "low", "close", "volume"]] for r in res_br if r["status"] == "success"]
    m_rows = sum(len(m) for m in minute)
    out["convert_frames"] = {"api": "InterpolatedToOHLCVConverter.convert_frames(frames) (spread_simulation; numpy's MT19937 draws the tapes on the host, "
                                    "in the reference's order, so a seeded run reproduces the reference's candles)",
                             "rows_in": rows, "candles_out": m_rows, "seconds": round(t_br, 4), "rate": rows / t_br, "rate_unit": "minute-rows/s"}
    rec = CandleReconstructor("5min")
    t_b, res_b = best(lambda: rec.reconstruct_batch(minute))
    out["reconstruct_batch"] = {"api": "CandleReconstructor('5min').reconstruct_batch(frames)", "rows_in": m_rows,
                                "candles_out": sum(len(r) for r in res_b if r is not None), "seconds": round(t_b, 4), "rate": m_rows / t_b,
                                "rate_unit": "minute-rows/s"}
    # the columnar entries: ONE Arrow table with every contract's rows in, one table out (tables.py) — no DataFrame per symbol
    import pyarrow as pa
    from iv_interpolation_b200 import tables as TB
    big = pa.Table.from_pandas(pd.concat(frames, ignore_index=True), preserve_index=False)
    t_t, (tab, _) = best(lambda: TB.interpolate_table(big))
    out["interpolate_table"] = {"api": "tables.interpolate_table(arrow_table): all contracts in one table, symbol as dictionary codes, "
                                       "result columns through pinned memory with one synchronisation", "rows_out": len(tab),
                                "seconds": round(t_t, 4), "rate": len(tab) / t_t, "rate_unit": "minute-rows/s"}
    t_pt, res_pt = best(lambda: TB.pipeline_table(big, seed=3))
    out["pipeline_table"] = {"api": "tables.pipeline_table(arrow_table): Task 1 -> bridge -> Task 2 chained on the device, the 5-min candle table back",
                             "seconds": round(t_pt, 4), "rate": rows / t_pt, "rate_unit": "minute-rows/s (Task-1 output rows)",
                             "candles_5min": len(res_pt["candles"])}
    fr = {f"c{i}": f for i, f in enumerate(frames)}
    t_p, res_p = best(lambda: PipelineBatch(rng=np.random.RandomState(2), batch_id=1).run(fr), reps=2)
    out["pipeline_batch"] = {"api": "PipelineBatch.run(frames): Task 1 -> bridge -> Task 2 (complete_pipeline.py's order), frames at every stage",
                             "seconds": round(t_p, 4), "rate": rows / t_p, "rate_unit": "minute-rows/s (Task-1 output rows)",
                             "candles_5min": sum(len(v) for v in res_p[0].values())}
    return out


def _dropin_ref_worker(job):
    kind, idx = job
    from oracle import reference_loader as RL
    fr = _DROPIN_REF["frames"]
    if kind == "a":
        t0 = time.perf_counter(); r = RL.reference_interpolate(fr[idx]); return (0 if r is None else len(r)), time.perf_counter() - t0
    m = _DROPIN_REF["minute"][idx % len(_DROPIN_REF["minute"])]
    t0 = time.perf_counter(); r = RL.reference_reconstruct(m, "5min"); return len(m), time.perf_counter() - t0


_DROPIN_REF = {}


def cpu_dropin_ref():
    """The reference's classes looped over frames of the same config-1 tape: one core, and (Task 1, Task 2) a
    ProcessPoolExecutor over all cores, one future per frame (batch_processor.py:234-239)."""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    from oracle import reference_loader as RL
    if not RL.available():
        return {"unavailable": "neither /root/reference nor oracle/_ref is present"}
    RL.quiet()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = 240
    frames = _dropin_frames(n)
    _DROPIN_REF["frames"] = frames
    RL.reference_interpolate(frames[0])
    t0 = time.perf_counter()
    outs = [RL.reference_interpolate(f) for f in frames[:96]]
    t_a = time.perf_counter() - t0
    rows_a = sum(len(o) for o in outs if o is not None)
    ok = [o for o in outs if o is not None]
    conv = RL.reference_converter()
    b_in = [o[["symbol", "date", "underlying_price", "mark_price", "index_price", "volume", "is_interpolated"]].rename(columns={"date": "timestamp"}) for o in ok[:40]]
    t0 = time.perf_counter()
    minute = [conv._generate_ohlcv_from_interpolated(b) for b in b_in]
    t_br = time.perf_counter() - t0
    rows_br = sum(len(b) for b in b_in)
    minute = [m[["symbol", "timestamp", "open", "high", "low", "close", "volume"]] for m in minute if m is not None]
    _DROPIN_REF["minute"] = minute
    RL.reference_reconstruct(minute[0], "5min")
    t0 = time.perf_counter()
    for k in range(96):
        RL.reference_reconstruct(minute[k % len(minute)], "5min")
    t_b = time.perf_counter() - t0
    rows_b = sum(len(minute[k % len(minute)]) for k in range(96))
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("fork")) as ex:
        list(ex.map(_dropin_ref_worker, [("a", 0)] * cores))
        t0 = time.perf_counter(); ra = list(ex.map(_dropin_ref_worker, [("a", i % n) for i in range(cores * 24)])); tp_a = time.perf_counter() - t0
        t0 = time.perf_counter(); rb = list(ex.map(_dropin_ref_worker, [("b", i) for i in range(cores * 24)])); tp_b = time.perf_counter() - t0
    return {"kind": "reference", "source": RL.where(), "cores": cores,
            "one_core": {"interpolate_symbol": {"rate": rows_a / t_a, "seconds_per_call": t_a / 96, "frames": 96},
                         "_generate_ohlcv_from_interpolated": {"rate": rows_br / t_br, "seconds_per_call": t_br / len(b_in), "frames": len(b_in)},
                         "reconstruct_symbol_candles": {"rate": rows_b / t_b, "seconds_per_call": t_b / 96, "frames": 96},
                         "pipeline": {"rate": 1.0 / (t_a / rows_a + t_br / rows_br + t_b / rows_b)}},
            "all_cores": {"interpolate_symbol": {"rate": sum(r[0] for r in ra) / tp_a, "frames": len(ra)},
                          "reconstruct_symbol_candles": {"rate": sum(r[0] for r in rb) / tp_b, "frames": len(rb)}},
            "rate_unit": "minute-rows/s", "sample": "frames of the same config-1 tape (tapes.make_contract_tape(0, n)): see `frames` per entry"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--n-t", type=int, default=0, help="stamps of the day (default: the full 23 400)")
    ap.add_argument("--ab-contracts", type=int, default=2400)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-kernels", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--workload", choices=("surface", "contracts", "dropin"), default="surface")
    ap.add_argument("--skip-dropin", action="store_true")
    ap.add_argument("--skip-stress", action="store_true")
    ap.add_argument("--dropin-contracts", type=int, default=DROPIN_CONTRACTS)
    ap.add_argument("--skip-contracts", action="store_true", help="leave the contracts object out of the surface line")
    ap.add_argument("--ct-underlyings", type=int, default=CT_UNDERLYINGS)
    ap.add_argument("--ct-windows", type=int, default=1, help="contracts workload: consecutive 20-day windows per underlying (12 ~ one trading year: BASELINE configs[4])")
    ap.add_argument("--ct-nan-frac", type=float, default=0.05, help="contracts workload: share of NaN IVs in the tape")
    ap.add_argument("--ct-drop-frac", type=float, default=0.0, help="contracts workload: share of hourly rows missing altogether (sparse replay tape)")
    ap.add_argument("--ct-two-step-bridge", action="store_true",
                    help="contracts workload: write the bridge's random tapes with k_random_tape, read them in the bridge kernel and gather the stamps behind it (the form before the draws moved into the kernel; same candles)")
    ap.add_argument("--ct-batch-rows", type=int, default=48_000_000, help="minute rows per batch of contracts (HBM working set)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: fewer than 3 warm-up steps breaks the timing rules; continuing as asked", file=sys.stderr)
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit

    def _emit(line: dict) -> None:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
