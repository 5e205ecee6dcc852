#!/usr/bin/env python
"""Summarise ncu output for profiles/: either a launch list (--metrics gpu__time_duration.sum CSV log)
or a full .ncu-rep (read through `ncu -i ... --page raw --csv`).  Usage:
    python tools/ncu_summary.py launches gpurun_out/x.csv  > profiles/x_launches.md
    python tools/ncu_summary.py full gpurun_out/x.ncu-rep  > profiles/x_full.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("smsp__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64cyc%"),
    ("smsp__cycles_active.avg", "cyc"),
    ("launch__grid_size", "grid"),
]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("rk::", "")


def to_unit(v, u):
    v = float(str(v).replace(",", ""))
    if u in ("ns", "nsecond"):
        return v / 1e3, "us"
    if u in ("us", "usecond"):
        return v, "us"
    if u in ("ms", "msecond"):
        return v * 1e3, "us"
    if u in ("s", "second"):
        return v * 1e6, "us"
    if u == "Kbyte":
        return v * 1e3, "B"
    if u == "Mbyte":
        return v * 1e6, "B"
    if u == "Gbyte":
        return v * 1e9, "B"
    if u == "byte":
        return v, "B"
    return v, u


def launches(path, traffic_json=None, config=None):
    """launch list (gpu__time_duration.sum per launch); when the same pass also holds dram__bytes_read.sum / dram__bytes_write.sum
    the table gets DRAM columns and, with `traffic_json config`, the per-kernel averages are merged into that JSON file
    (bench.py reads it for the `traffic` of its roofline object)."""
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        m = row.get("Metric Name")
        if m not in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
            continue
        v, _ = to_unit(row["Metric Value"], row["Metric Unit"])
        a = agg.setdefault(short(row["Kernel Name"]), [0, 0.0, 0.0, 0.0])
        if m == "gpu__time_duration.sum":
            a[0] += 1
            a[1] += v
        elif m == "dram__bytes_read.sum":
            a[2] += v
        else:
            a[3] += v
    tot = sum(v[1] for v in agg.values())
    has_dram = any(v[2] or v[3] for v in agg.values())
    print(f"# launch list summary of `{path}` (ncu --metrics gpu__time_duration.sum{', dram__bytes_read.sum, dram__bytes_write.sum' if has_dram else ''} --clock-control none; cold-cache, serialised: compare SHARES)\n")
    print("| kernel | launches | total us | avg us | share |" + (" DRAM read MB / launch | DRAM write MB / launch | DRAM GB/s |" if has_dram else "") + "\n|---|---:|---:|---:|---:|" + ("---:|---:|---:|" if has_dram else ""))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        extra = f" {v[2] / v[0] / 1e6:.1f} | {v[3] / v[0] / 1e6:.1f} | {(v[2] + v[3]) / v[1] / 1e3:.0f} |" if has_dram else ""
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |" + extra)
    print(f"\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
    if traffic_json and config and has_dram:
        import json, os
        t = json.load(open(traffic_json)) if os.path.exists(traffic_json) else {"source": "", "configs": {}}
        t["source"] = "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over the bench command (tools/gpu_call.sh r2f); per-launch averages"
        t.setdefault("configs", {})[config] = {k: {"launches": v[0], "dram_read_bytes_per_launch": v[2] / v[0], "dram_write_bytes_per_launch": v[3] / v[0],
                                                   "avg_us_under_ncu": v[1] / v[0]} for k, v in agg.items()}
        json.dump(t, open(traffic_json, "w"), indent=1)


def full(path):
    if path.endswith(".csv"):   # raw page already exported on the GPU box (ncu -i rep --page raw --csv)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of `{path}` (per launch; --clock-control none)\n")
    names = []
    for key, lab in KEYS:
        if key in col and lab not in names:
            names.append(lab)
    print("| # | kernel | " + " | ".join(names) + " |\n|---|---|" + "---:|" * len(names))
    for r in data:
        vals = {}
        for key, lab in KEYS:
            if key in col and lab not in vals:
                v, u = to_unit(r[col[key]], units[col[key]]) if r[col[key]] not in ("", "n/a") else (float("nan"), "")
                vals[lab] = f"{v:.3g}{'' if u in ('%', '', 'register/thread', 'cycle') else ' ' + u}" if lab not in ("time", "dram_rd", "dram_wr") else (
                    f"{v:.1f} us" if lab == "time" else f"{v / 1e6:.1f} MB")
        print(f"| {r[col['ID']]} | `{short(r[col['Kernel Name']])}` | " + " | ".join(vals[n] for n in names) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
