#!/usr/bin/env python
"""Per-source-line summary of an ncu report:  srcprof.py <report.ncu-rep> <kernel regex> [top N]
Wraps `ncu -i ... --page source --print-source cuda,sass --csv` (needs -lineinfo and --import-source on)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, hdr, items, seen = None, None, [], set()      # several launches of the kernel in a report: keep the first
for r in rows:

    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        key = (fname, d["Line No"])
        if key in seen:
            continue
        seen.add(key)
        try:
            items.append((int(d["Instructions Executed"]), int(d["# Samples"] or 0), fname, d["Line No"], r[1].strip()[:100],
                          float(d["Avg. Threads Executed"] or 0)))
        except ValueError:
            pass
tot = sum(i[0] for i in items) or 1
smp = sum(i[1] for i in items) or 1
print(f"total warp instructions {tot}, samples {smp}")
for n, s, f, l, src, thr in sorted(items, reverse=True)[:top]:
    print(f"{100*n/tot:5.1f}% inst {100*s/smp:5.1f}% smp thr={thr:4.1f} {f}:{l:>4} {src}")
