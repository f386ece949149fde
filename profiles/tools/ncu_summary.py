#!/usr/bin/env python
"""Key metrics of every kernel in an ncu report (first launch of each kernel name):
    ncu_summary.py <report.ncu-rep> <summary.json> [<traffic.json> <units.json>]
units.json maps a kernel-name prefix to {"units": N, "unit": "...", "algorithmic_bytes_per_unit": B} for the launch captured;
traffic.json then gets dram__bytes_read.sum + dram__bytes_write.sum per unit next to the algorithmic figure."""
import csv, json, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__grid_size", "launch__block_size"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short(name):
    m = re.search(r"(k_[A-Za-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    seen, res = set(), []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        k = short(d["Kernel Name"])
        if k in seen:
            continue
        seen.add(k)
        e = {"kernel": k}
        for key in KEYS:
            if key in d:
                e[key] = f"{d[key]} {units[hdr.index(key)]}".strip()
        res.append(e)
    json.dump(res, open(out, "w"), indent=1)
    if len(sys.argv) > 4:
        spec = json.load(open(sys.argv[4]))
        tr = {"_how": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full --clock-control none; " + spec.get("_how", "")}
        for e in res:
            for pref, u in spec.items():
                if pref.startswith("_") or not e["kernel"].startswith(pref) or pref in tr:
                    continue
                tot = 0.0
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v, un = e[key].split()
                    tot += float(v.replace(",", "")) * SCALE[un]
                tr[pref] = {"dram_bytes": int(tot), "units": u["units"], "unit": u["unit"], "bytes_per_unit": round(tot / u["units"], 2),
                            "algorithmic_bytes_per_unit": u["algorithmic_bytes_per_unit"]}
        json.dump(tr, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
