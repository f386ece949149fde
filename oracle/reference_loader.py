"""Import the reference's own hot-path classes.  TEST INFRASTRUCTURE ONLY.

Root, in this order: ``$IVX_REFERENCE_ROOT``; ``/root/reference`` (the build container — the modules are imported
from where they lie); ``oracle/_ref`` (the byte-for-byte copy ``oracle/make_ref.py`` makes of the path's four files:
git-ignored, it travels to the GPU box with the snapshot, so the reference's classes can be timed and cross-checked
there too).  Used by ``oracle/make_golden*.py`` (golden-vector generation), ``tests/test_oracle_vs_reference.py`` and
the CPU legs of ``bench.py`` (``cpu_baseline.kind: "reference"``).

Task 1 needs a shim in this image: reference src/interpolation/core.py:68 calls
``Series.fillna(method='ffill')``, an argument pandas 3 removed; the resulting TypeError is
swallowed by core.py:83-85 and every call returns None.  The shim maps
``fillna(method='ffill')`` to ``.ffill()``, which is what pandas<=2 did.
"""
from __future__ import annotations

import contextlib
import importlib
import logging
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_COPY = os.path.join(_HERE, "_ref")


def _has_path(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "src", "interpolation", "core.py"))


def _resolve_root() -> str:
    env = os.environ.get("IVX_REFERENCE_ROOT")
    if env:
        return env
    return "/root/reference" if _has_path("/root/reference") else _REF_COPY


REFERENCE_ROOT = _resolve_root()


def available() -> bool:
    return _has_path(REFERENCE_ROOT)


def where() -> str:
    """'reference' = imported from /root/reference where it lies; '_ref' = the untouched copy under oracle/_ref."""
    return "_ref" if os.path.abspath(REFERENCE_ROOT) == os.path.abspath(_REF_COPY) else "reference"


@contextlib.contextmanager
def fillna_method_shim():
    import pandas as pd

    orig = pd.Series.fillna

    def fillna(self, value=None, *args, method=None, **kw):
        if method is not None:
            if method in ("ffill", "pad"):
                return self.ffill()
            if method in ("bfill", "backfill"):
                return self.bfill()
            raise ValueError(method)
        return orig(self, value, *args, **kw)

    pd.Series.fillna = fillna
    try:
        yield
    finally:
        pd.Series.fillna = orig


def _import(modname: str):
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    for name in ("interpolation", "candle_reconstruction", "data_bridge"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "") or getattr(mod, "__path__", [""])[0]).startswith(REFERENCE_ROOT):
            raise RuntimeError(f"{name} already imported from elsewhere: {mod}")
    return importlib.import_module(modname)


def quiet():
    logging.getLogger("interpolation.core").setLevel(logging.CRITICAL)
    logging.getLogger("candle_reconstruction.core").setLevel(logging.CRITICAL)
    logging.getLogger("data_bridge.ohlcv_converter").setLevel(logging.CRITICAL)


def load_interpolator_cls():
    quiet()
    return _import("interpolation.core").IVInterpolator


def load_reconstructor_cls():
    quiet()
    return _import("candle_reconstruction.core").CandleReconstructor


def load_greeks_cls():
    return _import("interpolation.greeks").BlackScholesGreeks


def reference_interpolate(df, method: str = "linear", min_points: int = 10):
    """IVInterpolator(method, min_points).interpolate_symbol(df) under the pandas-3 shim."""
    cls = load_interpolator_cls()
    with fillna_method_shim():
        return cls(method=method, min_points=min_points).interpolate_symbol(df)


def reference_reconstruct(df, target_frequency: str = "5min"):
    cls = load_reconstructor_cls()
    return cls(target_frequency).reconstruct_symbol_candles(df)


def load_converter_cls():
    quiet()
    return _import("data_bridge.ohlcv_converter").InterpolatedToOHLCVConverter


def reference_converter(strategy: str = "spread_simulation", params=None):
    """InterpolatedToOHLCVConverter(None, config) with the reference's default bridge configuration (config.py:57-75)."""
    import types
    cfg = types.SimpleNamespace(data_bridge=types.SimpleNamespace(
        conversion_strategy=strategy, spread_method="adaptive", enable_quality_checks=True,
        spread_parameters=params if params is not None else {"base_spread_percent": 0.002, "volatility_factor": 1.5}))
    return load_converter_cls()(None, cfg)
