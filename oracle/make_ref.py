#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: the reference's OWN hot-path modules, byte for byte.  TEST INFRASTRUCTURE ONLY.

The reference is pure Python; "building" it means making its hot-path modules travel to the GPU box, where
/root/reference does not exist.  This script copies — untouched, checked by sha256 — the four files SURVEY.md
section 8(a) puts on the path from where they lie under /root/reference into ``oracle/_ref/src/...`` (same package
layout, so ``oracle/reference_loader.py`` imports them the same way it imports the originals).  ``oracle/_ref/`` is
git-ignored (no reference source ever enters the history) but not gpurun-ignored.  ``__graft_entry__.build()`` runs
this recipe whenever /root/reference is present; on the GPU box the prebuilt directory is used as is.

    python oracle/make_ref.py          # refresh oracle/_ref from /root/reference
    python oracle/make_ref.py --check  # verify oracle/_ref against its manifest (and the originals when present)

What uses it: ``bench.py --impl reference`` / the ``cpu_baseline`` legs (kind "reference": the reference's classes
timed on the box's host cores), ``tests/test_oracle_vs_reference.py`` (oracle restatements vs the reference, bit
for bit).  Never the product: ``iv_interpolation_b200`` does not import ``oracle``.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("IVX_REFERENCE_SRC", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")

# the files of SURVEY.md 8(a) + 8(f-2); nothing else of the reference is needed to run the path
FILES = (
    "src/interpolation/__init__.py",
    "src/interpolation/core.py",               # IVInterpolator.interpolate_symbol            (kernel A)
    "src/interpolation/greeks.py",             # BlackScholesGreeks                            (BS conventions)
    "src/candle_reconstruction/core.py",       # CandleReconstructor.reconstruct_symbol_candles (kernel B)
    "src/data_bridge/ohlcv_converter.py",      # InterpolatedToOHLCVConverter                  (the step between A and B)
)


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def make() -> dict:
    if not os.path.isdir(SRC_ROOT):
        raise SystemExit(f"make_ref: {SRC_ROOT} is not present; oracle/_ref can only be made in the build container")
    manifest = {"source_root": SRC_ROOT, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest["files"][rel] = _sha(dst)
        if _sha(src) != manifest["files"][rel]:
            raise SystemExit(f"make_ref: copy of {rel} differs from its source")
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


def check() -> bool:
    mpath = os.path.join(DST_ROOT, "MANIFEST.json")
    if not os.path.isfile(mpath):
        return False
    with open(mpath) as f:
        manifest = json.load(f)
    for rel, digest in manifest["files"].items():
        p = os.path.join(DST_ROOT, rel)
        if not os.path.isfile(p) or _sha(p) != digest:
            return False
        src = os.path.join(SRC_ROOT, rel)
        if os.path.isfile(src) and _sha(src) != digest:
            return False
    return set(manifest["files"]) == set(FILES)


if __name__ == "__main__":
    if "--check" in sys.argv[1:]:
        ok = check()
        print("oracle/_ref", "ok" if ok else "MISSING OR STALE")
        sys.exit(0 if ok else 1)
    m = make()
    print(f"oracle/_ref: {len(m['files'])} files copied untouched from {SRC_ROOT}")
