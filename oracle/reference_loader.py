"""Import the reference's own hot-path classes from /root/reference.  TEST INFRASTRUCTURE ONLY.

Used by ``oracle/make_golden.py`` (golden-vector generation) and by the container-only
cross-checks in ``tests/test_oracle_vs_reference.py`` (skipped where /root/reference is
absent, i.e. on the GPU box).  Nothing here is copied from the reference: the modules are
imported from where they lie.

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

REFERENCE_ROOT = os.environ.get("IVX_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "interpolation", "core.py"))


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
    for name in ("interpolation", "candle_reconstruction"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "") or getattr(mod, "__path__", [""])[0]).startswith(REFERENCE_ROOT):
            raise RuntimeError(f"{name} already imported from elsewhere: {mod}")
    return importlib.import_module(modname)


def quiet():
    logging.getLogger("interpolation.core").setLevel(logging.CRITICAL)
    logging.getLogger("candle_reconstruction.core").setLevel(logging.CRITICAL)


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
