"""CPU tests of the host side: the C ABI surface (header <-> library <-> ctypes table), the error contract of the
drop-in classes without a GPU, and the world_size-2 sharding/gather logic over gloo."""
import ctypes
import os
import re
import socket

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ivx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ivx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from iv_interpolation_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} is declared in include/ivx.h but not exported by libivx.so"
    missing = [n for n in names if n not in _lib.SIGNATURES]
    assert not missing, f"ctypes table lacks {missing}"
    L = _lib.lib()                      # loads and binds every signature; no compute call is made
    assert L.ivx_version() == 100
    info = L.ivx_build_info().decode()
    for k in ("k_interp_plan", "k_interp_exec", "k_ohlcv_reduce", "k_bs_implied_vol", "k_spline_fit_eval"):
        assert k in info


def test_library_is_sm100a_only():
    """The shipped binary carries sm_100a SASS and nothing else (no multi-arch fat binary)."""
    import subprocess
    from iv_interpolation_b200 import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_host_floor_div_matches_python():
    from iv_interpolation_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(0)
    for d in (1, 2, 3, 5, 7, 60, 300, 3600, 60_000_000, 300_000_000, 7 * 60_000_000_000, 2 ** 31, 2 ** 40 + 1):
        ts = np.concatenate([rng.integers(-2 ** 62, 2 ** 62, 200), np.array([0, 1, -1, d, -d, d - 1, -d + 1, 2 ** 62, -2 ** 62])])
        for t in ts:
            assert L.ivx_host_floor_div(int(t), int(d)) == int(t) // int(d), (t, d)


def test_argument_errors_come_back_as_codes_not_crashes():
    from iv_interpolation_b200 import _lib
    L = _lib.lib()
    rc = L.ivx_bs_implied_vol(-1, None, None, None, None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"negative" in L.ivx_last_error_string()
    rc = L.ivx_ohlcv_reduce(1, None, 4, None, None, None, None, None, None, None, 0, 5, 1, 0, None, None, None, None, None,
                            None, None, None, None, None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"bucket_ticks" in L.ivx_last_error_string()
    assert L.ivx_spline_strided(10, 65, None, None, 4, None, 1, None, None) == -4       # IVX_ERR_UNSUPPORTED
    assert L.ivx_ohlcv_workspace_bytes(10, 5000) > 0 and L.ivx_interp_workspace_bytes(-1, 0, 0, 0) == -1


def test_dropin_contract_without_gpu():
    """The reference's early-outs do not need the device; anything that does must fail loudly inside and
    surface as None (the reference's never-raise contract), never as a silent CPU result."""
    import torch
    from iv_interpolation_b200.candle_reconstruction.core import CandleReconstructor
    from iv_interpolation_b200.interpolation.core import IVInterpolator
    assert IVInterpolator().interpolate_symbol(pd.DataFrame({"date": [], "iv": []})) is None          # < min_points
    with pytest.raises(ValueError):
        CandleReconstructor("1h")                                                                        # candle core.py:34
    assert CandleReconstructor("15m").frequency_minutes == 15
    assert CandleReconstructor("5min").reconstruct_symbol_candles(pd.DataFrame()) is None               # empty
    if not torch.cuda.is_available():
        ts = pd.date_range("2023-03-15", periods=30, freq="1min")
        df = pd.DataFrame({"symbol": "s", "timestamp": ts, "open": 1.0, "high": 2.0, "low": 0.5, "close": 1.5, "volume": 1.0})
        assert CandleReconstructor("5min").reconstruct_symbol_candles(df) is None                       # no device -> error -> None
        from iv_interpolation_b200._lib import IvxError
        with pytest.raises(IvxError):
            CandleReconstructor("5min").reconstruct_batch([df])


def test_bridge_contract_without_gpu():
    """The bridge class keeps the reference's never-raise contract (ohlcv_converter.py:141-142, :186-188) and, without a
    device, fails inside (logged) instead of producing candles on the CPU; its tape draws still follow numpy's generator."""
    import types
    import torch
    from iv_interpolation_b200.data_bridge import ohlcv_converter as OC
    cfg = types.SimpleNamespace(data_bridge=types.SimpleNamespace(conversion_strategy="price_midpoint", spread_method="adaptive",
                                                                  enable_quality_checks=True, spread_parameters=None))
    conv = OC.InterpolatedToOHLCVConverter(None, cfg)
    assert conv._generate_ohlcv_from_interpolated(pd.DataFrame()) is None
    assert OC.canonical_strategy("price_midpoint") == "simple_spread"          # the config comment's name is not a strategy
    ts = pd.date_range("2024-01-01", periods=10, freq="1min")
    df = pd.DataFrame({"symbol": "s", "timestamp": ts, "underlying_price": 100.0, "mark_price": np.nan, "volume": 1.0})
    assert conv._select_price_column(df) == "underlying_price"
    assert conv._select_price_column(df.assign(underlying_price=np.nan)) == "underlying_price"      # fallback: first present
    with pytest.raises(ValueError):
        conv._select_price_column(df[["symbol", "timestamp", "volume"]])
    ok = np.array([True, False, True, True])
    need = np.array([False, False, True, False])
    a, b = np.random.RandomState(4), np.random.RandomState(4)
    u, e, z = OC.draw_tapes(a, "spread_simulation", ok, need)
    want = [b.uniform(0, 1) for _ in range(5)] + [b.uniform(0, 1) for _ in range(5)]
    we = b.exponential(1.0)
    want += [b.uniform(0, 1) for _ in range(5)]
    assert np.array_equal(u, want) and np.array_equal(e, [we]) and len(z) == 0 and a.random_sample() == b.random_sample()
    if not torch.cuda.is_available():
        assert conv._generate_ohlcv_from_interpolated(df) is None                                   # no device -> error -> None
        assert all(r["status"] == "error" for r in conv.convert_frames([df]))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "iv_interpolation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


# ---- sharding ---------------------------------------------------------------------------------------------
def test_lpt_partition_properties():
    from iv_interpolation_b200 import sharding as SH
    rng = np.random.default_rng(1)
    costs = rng.integers(1, 100_000, 1200).astype(float)
    for world in (1, 2, 4, 8):
        shards = SH.lpt_partition(costs, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(costs)))
        assert all(s == sorted(s) for s in shards)
        assert SH.imbalance(costs, shards) < 1.01
    assert SH.lpt_partition([5.0], 4) == [[0], [], [], []]
    assert SH.lpt_partition([], 2) == [[], []]
    off = np.array([0, 3, 3, 5]); ts = np.array([0, 60, 600, 1000, 1120])
    assert SH.contract_costs(off, ts, 60).tolist() == [11.0, 0.0, 3.0]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from iv_interpolation_b200 import sharding as SH
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        costs = [float(c) for c in (7, 3, 9, 1, 4, 4, 8, 2)]
        mine = SH.my_units(costs)
        # each rank "processes" its units: one candle row per unit of cost
        ts = torch.tensor([u * 1000 + k for u in mine for k in range(int(costs[u]))], dtype=torch.int64)
        close = ts.to(torch.float64) * 0.5
        table = SH.gather_tables({"ts": ts, "close": close, "unit": torch.tensor([u for u in mine for _ in range(int(costs[u]))], dtype=torch.int32)})
        if rank == 0:
            q.put((mine, {k: v.tolist() for k, v in table.items()}))
        else:
            assert table is None
            q.put((mine, None))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    units = sorted(u for mine, _ in got for u in mine)
    assert units == list(range(8))                                   # every unit on exactly one rank
    table = next(t for _, t in got if t is not None)
    costs = [7, 3, 9, 1, 4, 4, 8, 2]
    assert len(table["ts"]) == sum(costs)
    assert table["close"] == [t * 0.5 for t in table["ts"]]
    assert sorted(set(table["unit"])) == list(range(8))
    # rank order, then unit order inside a rank
    assert table["_rank"] == sorted(table["_rank"])
    from iv_interpolation_b200 import sharding as SH
    shards = SH.lpt_partition([float(c) for c in costs], 2)
    want_units = [u for s in shards for u in s for _ in range(costs[u])]
    assert table["unit"] == want_units


def test_columnar_round_trip_at_db_precision(tmp_path):
    """Frames -> Arrow/Parquet with the reference's DECIMAL(p, s) columns -> per-symbol frames ordered by time."""
    import pyarrow as pa
    from iv_interpolation_b200 import columnar as CO
    ts = pd.date_range("2023-03-15", periods=6, freq="1min")
    f1 = pd.DataFrame({"date": ts, "symbol": "btc-17mar23-24500-c", "iv": [0.123456789012, 0.5, 0.25, 1.0, 0.3333333349, 0.75],
                       "underlying_price": 25000.12345, "time_to_maturity": 0.0123456789, "strike": 24500.0, "callput": "c",
                       "interest_rate": 0.05, "mark_price": 0.0312, "index_price": 25001.5, "volume": 1.123456789123,
                       "quote_volume": 3.5, "record_time": "rt", "is_interpolated": False, "batch_id": 7})
    f2 = f1.assign(symbol="btc-17mar23-25000-p", callput="p").iloc[::-1]
    t = CO.interpolated_to_table({"a": f1, "b": f2, "none": None})
    assert t.schema.field("iv").type == pa.decimal128(12, 8) and t.schema.field("underlying_price").type == pa.decimal128(15, 4)
    assert t.schema.field("volume").type == pa.decimal128(20, 8) and t.num_rows == 12
    assert t.column_names[:3] == ["symbol", "date", "record_time"]
    path = str(tmp_path / "interp.parquet")
    CO.write_parquet(t, path)
    back = CO.read_symbol_frames(path, "date")
    assert set(back) == {"btc-17mar23-24500-c", "btc-17mar23-25000-p"}
    b1, b2 = back["btc-17mar23-24500-c"], back["btc-17mar23-25000-p"]
    assert np.allclose(b1["iv"], np.round(f1["iv"], 8), atol=0, rtol=0) and b1["underlying_price"].iloc[0] == 25000.1234 + 0.0
    assert list(b2["date"]) == list(ts)                                   # re-ordered by time
    assert b1["volume"].iloc[0] == 1.12345679
    c = pd.DataFrame({"symbol": "s", "timestamp": ts[:2], "open": [1.00004, 2.0], "high": [2.5, 3.0], "low": [0.5, 1.0],
                      "close": [1.5, 2.5], "volume": [10.0, 0.123456785], "frequency": "5min", "source_candles": 5,
                      "created_at": ts[:2]})
    ct = CO.candles_to_table({"s": c})
    assert ct.schema.field("open").type == pa.decimal128(15, 4) and ct.column_names == list(CO.CANDLE_ORDER[:9]) + ["created_at"]
    assert float(ct.column("open")[0].as_py()) == 1.0


def test_contracts_workload_plan_is_fixed_and_balanced():
    """bench.py's contracts workload (BASELINE.json configs[2]): the tape plan does not depend on the number of ranks, every
    underlying lands on exactly one rank, and longest-first dealing keeps max/mean shard load within a few percent at 8 ranks."""
    import bench
    from iv_interpolation_b200 import sharding as SH
    n_c, costs = bench._ct_plan(100)
    n_c2, _ = bench._ct_plan(100)
    assert (n_c == n_c2).all() and 900 <= n_c.min() and n_c.max() <= 1500 and abs(n_c.mean() - 1200) < 60
    for world in (1, 2, 4, 8):
        shards = SH.lpt_partition(costs, world)
        assert sorted(u for s in shards for u in s) == list(range(100))
        assert SH.imbalance(costs, shards) < 1.03
    shard = bench._ct_build_shard([3, 5], n_c, pin=False)
    assert shard.n_contracts == n_c[3] + n_c[5] and int(shard.off[-1]) == shard.n_contracts * bench.CT_ROWS
    assert shard.ids[0] == n_c[:3].sum() and shard.ids[-1] == n_c[:6].sum() - 1 and (shard.out_rows == bench.CT_OUT_ROWS).all()
