#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep captured with --import-source on (and -lineinfo).
usage: ncu_lines.py rep kernel-regex [top-n] [substring the demangled function name must contain]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
need = sys.argv[4] if len(sys.argv) > 4 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict()
hdr = None
cur_file = None
cur_fn = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name":
        cur_fn = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if need and need not in cur_fn:
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    d = dict(zip(hdr[2:], r[2:]))   # sass-level columns (the 2nd "Source" is the SASS text)
    key = (cur_file, int(r[0]), r[1].strip()[:110])
    a = agg.setdefault(key, [0, 0, 0, 0])
    def num(x):
        try: return float(x)
        except: return 0.0
    a[0] += num(d.get("# Samples")); a[1] += num(d.get("Instructions Executed")); a[2] += num(d.get("stall_long_sb")); a[3] += num(d.get("L2 Theoretical Sectors Global"))
tot_s = sum(a[0] for a in agg.values()) or 1; tot_i = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
print(f"{'samp%':>6} {'inst%':>6} {'longsb':>7} {'sectors':>10}  file:line  source")
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot_s:6.1f} {100*a[1]/tot_i:6.1f} {a[2]:7.0f} {a[3]:10.0f}  {f}:{ln}  {src}")
