"""ctypes binding of libivx.so (the C ABI declared in include/ivx.h).

There is NO fallback: if the shared library is missing this module raises, and every compute
entry point of the package goes through it.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IVX_LIB_PATH") or os.path.join(_HERE, "csrc", "libivx.so")   # IVX_LIB_PATH: profiling builds

# error / status codes (include/ivx.h)
IVX_OK = 0
SYM_OK, SYM_TOO_FEW, SYM_SPAN, SYM_TOO_LONG, SYM_EMPTY_OUT, SYM_NO_DATES = 0, 1, 2, 3, 4, 5
FLAG_IN_BAD_OHLC, FLAG_IN_NEG_VOLUME, FLAG_IN_NAN = 1, 2, 4
FLAG_OUT_BAD_OHLC, FLAG_OUT_NEG_VOLUME, FLAG_OUT_NAN = 8, 16, 32
FLAG_UNSORTED, FLAG_NONFINITE = 64, 128
MAX_FILL_COLS, MAX_IDX_COLS = 8, 4
BRIDGE_STRATEGY = {"spread_simulation": 0, "price_as_midpoint": 1, "trend_following": 2, "simple_spread": 3, "pipeline": 4}
BRIDGE_U_PER_ROW = {0: 5, 1: 2, 2: 0, 3: 1, 4: 4}
BRIDGE_RECOMPUTE = 0x100          # or-ed into the strategy code: per-symbol kernel after the segment kernel whatever the check says
Q_HIGH_LT_LOW, Q_HIGH_LT_OC, Q_LOW_GT_OC, Q_SPREAD, Q_NONPOSITIVE = 1, 2, 4, 8, 16


class IvxError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None

_i64, _i32, _p = C.c_int64, C.c_int32, C.c_void_p
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/ivx.h one to one
SIGNATURES = {
    "ivx_version": (C.c_int, []),
    "ivx_last_error_string": (C.c_char_p, []),
    "ivx_build_info": (C.c_char_p, []),
    "ivx_ohlcv_workspace_bytes": (_i64, [_i64, _i64]),
    "ivx_ohlcv_reduce": (C.c_int, [_i64, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i64, _p,
                                   _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ivx_time_buckets": (C.c_int, [_i64, _p, _i64, _p, _p, _p, _p, _p]),
    "ivx_ohlc_time_major": (C.c_int, [_i64, _i64, _p, _i64, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "ivx_interp_workspace_bytes": (_i64, [_i64, _i64, _i32, _i32]),
    "ivx_interp_plan": (C.c_int, [_i64, _p, _i64, _p, _pp, _i32, _pp, _i32, _pp, _i64, _i64, _i64, _i32,
                                  _p, _p, _p, _p, _p, _i64, _p]),
    "ivx_interp_exec": (C.c_int, [_i64, _p, _i64, _p, _i64, _i64, _i32, _i32, _p, _pp, _pp, _pp, _p, _p, _p,
                                  _p, _p, _i64, _p]),
    "ivx_bs_implied_vol_workspace_bytes": (_i64, [_i64]),
    "ivx_bs_implied_vol": (C.c_int, [_i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ivx_bs_price": (C.c_int, [_i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ivx_bs_greeks": (C.c_int, [_i64, _p, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p]),
    "ivx_spline_workspace_bytes": (_i64, [_i64, _i64]),
    "ivx_spline_fit_eval": (C.c_int, [_i64, _p, _i64, _p, _p, _i64, _p, _i32, _p, _p, _i64, _p]),
    "ivx_spline_fit_eval_xoff": (C.c_int, [_i64, _p, _p, _i64, _p, _p, _i64, _p, _i32, _p, _p]),
    "ivx_spline_ladder_table_len": (_i64, [_i64, _i64]),
    "ivx_spline_ladder_factor": (C.c_int, [_i64, _p, _i64, _p, _p, _p, _p, _p]),
    "ivx_spline_fit_eval_ladder": (C.c_int, [_i64, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _i64, _p, _i32, _p, _p]),
    "ivx_spline_fit_eval_ladder_verify": (C.c_int, [_i64, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _i64, _p, _i32, _p, _p]),
    "ivx_bs_chain_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "ivx_bs_implied_vol_chain": (C.c_int, [_i64, _i64, _i64, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ivx_spline_band_len": (_i64, [_i64, _i64]),
    "ivx_spline_band_frag_len": (_i64, [_i64]),
    "ivx_spline_band_oct_len": (_i64, [_i64]),
    "ivx_spline_band_fill_len": (_i64, [_i64]),
    "ivx_spline_band_build": (C.c_int, [_i64, _p, _p, _i64, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ivx_spline_fit_eval_banded": (C.c_int, [_i64, _p, _p, _p, _p, _p, _i64, _i64, _p, _i64, _p, _p, _p, _p, _p, _p, _i64, _p, _i32, _p, _p, _p,
                                             _p, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p, _p]),
    "ivx_bs_implied_vol_ladder_workspace_bytes": (_i64, [_i64, _i64]),
    "ivx_bs_implied_vol_ladder": (C.c_int, [_i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ivx_bs_implied_vol_chain_ticks": (C.c_int, [_i64, _i64, _i64, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ivx_spline_strided": (C.c_int, [_i64, _i32, _p, _p, _i32, _p, _i32, _p, _p]),
    "ivx_spline_strided_batched": (C.c_int, [_i64, _i64, _i32, _p, _i64, _i64, _p, _i32, _p, _i32, _p, _p]),
    "ivx_bridge_n_segments": (_i64, [_i64]),
    "ivx_bridge_workspace_bytes": (_i64, [_i64]),
    "ivx_bridge_quality": (C.c_int, [_i64, _p, _p, _p, _p, _p, _p, _p]),
    "ivx_bridge_plan": (C.c_int, [_i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ivx_bridge_ohlcv": (C.c_int, [_i64, _i64, _p, _p, _p, _i32, C.c_double, C.c_double, _p, _p, _p, _p, _p, _p, _p,
                                   _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ivx_bridge_ohlcv_philox": (C.c_int, [_i64, _i64, _p, _p, _p, _p, _i32, C.c_double, C.c_double, C.c_uint64, C.c_uint32, C.c_uint32,
                                          _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "ivx_exclusive_scan_i64": (C.c_int, [_i64, _p, _p, _p]),
    "ivx_gather_i64": (C.c_int, [_i64, _p, _i64, _p, _p, _p]),
    "ivx_gather_f64": (C.c_int, [_i64, _p, _i64, _p, _p, _p]),
    "ivx_read_back_i64": (C.c_int, [_i32, _p, C.POINTER(C.c_int64), _p]),
    "ivx_random_tape": (C.c_int, [_i64, C.c_uint64, C.c_uint32, _i32, _p, _p]),
    "ivx_probe_copy": (C.c_int, [_i64, _p, _p, _p]),
    "ivx_probe_fp64_fma": (C.c_int, [_i64, _i32, _p, _p]),
    "ivx_probe_write_streams": (C.c_int, [_i64, _p, _i64, _p]),
    "ivx_host_floor_div": (_i64, [_i64, _i64]),
}


def lib() -> C.CDLL:
    """Load libivx.so once.  Raises IvxError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise IvxError(
                f"{LIB_PATH} is missing: build it with `make -C iv_interpolation_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here = header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().ivx_last_error_string().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != IVX_OK:
        raise IvxError(f"{what or 'libivx'} failed with code {rc}: {last_error()}")


def ptr_array(ptrs):
    """Host array of device pointers (``const double* const*`` arguments)."""
    arr = (C.c_void_p * max(1, len(ptrs)))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
