#!/usr/bin/env python
"""Instruction mix of one kernel from `ncu -i X.ncu-rep --page source --csv --kernel-name K` (stdin or file)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = 0
ops, samp = collections.Counter(), collections.Counter()
n_static = 0
for r in rows[hi + 1:]:
    if len(r) <= ia or not r[ia].isdigit():
        continue
    n = int(r[ia]); tot += n; n_static += 1
    toks = [o for o in r[isrc].split() if not o.startswith("@")]
    op = toks[0].split(".")[0] if toks else "?"
    ops[op] += n; samp[op] += int(r[ist]) if r[ist].isdigit() else 0
print("total warp-instructions", tot, "static", n_static)
for o, n in ops.most_common(30):
    print(f"{o:10s} {n / 1e6:9.1f}M {100 * n / tot:5.1f}%  stall samples {samp[o]}")
