"""(De)serialise small DataFrames into flat npz dictionaries.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import json
from typing import Dict, Optional

import numpy as np
import pandas as pd


def frame_to_dict(prefix: str, df: Optional[pd.DataFrame], out: Dict[str, np.ndarray]) -> None:
    """Store ``df`` under ``prefix``; ``None`` is recorded as such."""
    if df is None:
        out[prefix + "/none"] = np.array(1)
        return
    meta = {"columns": [], "index": "range" if isinstance(df.index, pd.RangeIndex) else "array"}
    out[prefix + "/index"] = np.asarray(df.index, dtype=np.int64)
    for i, c in enumerate(df.columns):
        col = df[c]
        key = f"{prefix}/c{i}"
        dt = col.dtype
        if isinstance(dt, np.dtype) and dt.kind == "M":
            unit = np.datetime_data(dt)[0]
            out[key] = col.to_numpy().view(np.int64)
            meta["columns"].append([str(c), "M", unit])
        elif isinstance(dt, pd.DatetimeTZDtype):
            out[key] = col.dt.tz_convert("UTC").dt.tz_localize(None).to_numpy().view(np.int64)
            meta["columns"].append([str(c), "Mtz", dt.unit, str(dt.tz)])
        elif isinstance(dt, np.dtype) and dt.kind == "f":
            out[key] = col.to_numpy(dtype=np.float64)
            meta["columns"].append([str(c), "f"])
        elif isinstance(dt, np.dtype) and dt.kind == "b":
            out[key] = col.to_numpy(dtype=bool)
            meta["columns"].append([str(c), "b"])
        elif isinstance(dt, np.dtype) and dt.kind in "iu":
            out[key] = col.to_numpy(dtype=np.int64)
            meta["columns"].append([str(c), "i"])
        else:                                   # strings / objects: value + null mask
            mask = col.isna().to_numpy(dtype=bool)
            vals = np.array(["" if m else str(v) for v, m in zip(col.tolist(), mask)], dtype=np.str_)
            out[key] = vals
            out[key + "/null"] = mask
            meta["columns"].append([str(c), "U"])
    out[prefix + "/meta"] = np.array(json.dumps(meta))


def dict_to_frame(prefix: str, d) -> Optional[pd.DataFrame]:
    if prefix + "/none" in d:
        return None
    meta = json.loads(str(d[prefix + "/meta"]))
    cols = {}
    for i, spec in enumerate(meta["columns"]):
        name, kind = spec[0], spec[1]
        a = d[f"{prefix}/c{i}"]
        if kind == "M":
            cols[name] = a.view(f"datetime64[{spec[2]}]")
        elif kind == "Mtz":
            cols[name] = pd.Series(a.view(f"datetime64[{spec[2]}]")).dt.tz_localize("UTC").dt.tz_convert(spec[3])
        elif kind == "U":
            null = d[f"{prefix}/c{i}/null"]
            vals = a.astype(object)
            vals[null] = None
            cols[name] = pd.Series(vals, dtype="str" if not null.any() else object)
        else:
            cols[name] = a
    idx = d[prefix + "/index"]
    df = pd.DataFrame(cols)
    if meta["index"] == "range" and len(idx):
        df.index = pd.RangeIndex(int(idx[0]), int(idx[0]) + len(idx))
    else:
        df.index = pd.Index(idx)
    return df


def assert_frames_match(ref: Optional[pd.DataFrame], got: Optional[pd.DataFrame], *, ignore=("created_at",),
                        float_atol: float = 0.0) -> None:
    """Bit-exact comparison (NaN==NaN) of values, column order and index; string columns compare
    by value and null-ness (their dtype flavour — ``str`` vs ``object`` — is not part of parity)."""
    if ref is None or got is None:
        assert ref is None and got is None, f"None mismatch: reference {'None' if ref is None else 'frame'}, got {'None' if got is None else 'frame'}"
        return
    rc = [c for c in ref.columns if c not in ignore]
    gc = [c for c in got.columns if c not in ignore]
    assert rc == gc, f"columns differ: {rc} vs {gc}"
    assert len(ref) == len(got), f"row count differs: {len(ref)} vs {len(got)}"
    assert np.array_equal(np.asarray(ref.index), np.asarray(got.index)), "index differs"
    for c in rc:
        a, b = ref[c], got[c]
        ka = a.dtype.kind if isinstance(a.dtype, np.dtype) else "O"
        kb = b.dtype.kind if isinstance(b.dtype, np.dtype) else "O"
        if ka == "f" or kb == "f":
            x = a.to_numpy(dtype=np.float64)
            y = b.to_numpy(dtype=np.float64)
            if float_atol == 0.0:
                same = (x == y) | (np.isnan(x) & np.isnan(y))
            else:
                same = (np.abs(x - y) <= float_atol) | (np.isnan(x) & np.isnan(y)) | (x == y)
            assert same.all(), f"column {c!r}: {int((~same).sum())} of {len(x)} values differ, first at {int(np.argmin(same))}: {x[np.argmin(same)]!r} vs {y[np.argmin(same)]!r}"
        elif ka == "M" or kb == "M" or isinstance(a.dtype, pd.DatetimeTZDtype):
            assert str(a.dtype) == str(b.dtype), f"column {c!r}: dtype {a.dtype} vs {b.dtype}"
            assert (a.astype("int64").to_numpy() == b.astype("int64").to_numpy()).all(), f"column {c!r} differs"
        elif ka in "biu" and kb in "biu":
            assert ka == kb and (a.to_numpy() == b.to_numpy()).all(), f"column {c!r} differs"
        else:
            na, nb = a.isna().to_numpy(), b.isna().to_numpy()
            assert (na == nb).all(), f"column {c!r}: null masks differ"
            va = np.array([str(v) for v in a[~na].tolist()])
            vb = np.array([str(v) for v in b[~nb].tolist()])
            assert (va == vb).all(), f"column {c!r}: values differ"
