#!/usr/bin/env python
"""Static SASS evidence for profiles/: per kernel of librheo_b200.so, the counts of the opcodes that show what the code is —
sm_100a cubin, TMA bulk copies (UBLKCP), mbarrier (SYNCS), programmatic dependent launch (ACQBULK / PREEXIT), FP64 math
(DFMA / DADD / DMUL / MUFU.RCP64H), shared memory, warp shuffles, barriers, atomics.  No tensor-core opcode is expected: nothing
on this path is a dense contraction.      usage: python tools/sass_static.py [lib.so] > profiles/r2_sass_summary.md"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parent.parent / "rheotool_b200" / "librheo_b200.so"
elf = subprocess.run(["cuobjdump", "-lelf", str(lib)], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
WATCH = ["UBLKCP", "SYNCS", "ACQBULK", "PREEXIT", "DFMA", "DADD", "DMUL", "MUFU", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "WARPSYNC", "ATOM", "RED", "HMMA", "UTCMMA", "UTMALDG"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("rk::", "").replace("(int)", "")
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["total"] += 1
        for w in WATCH:
            if op.startswith(w):
                per[cur][w] += 1
print(f"# SASS opcode summary of {lib.name} (static counts; `cuobjdump -sass`, tools/sass_static.py)\n")
print("ELF images: " + ", ".join(sorted(set(re.findall(r"sm_\w+", elf)))) + "\n")
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("Whole library: " + ", ".join(f"{w} {tot[w]}" for w in WATCH if tot[w]) + f", instructions {tot['total']}\n")
print("| kernel | instr | " + " | ".join(WATCH[:17]) + " |")
print("|---|---|" + "---|" * 17)
for k, c in per.items():
    if c["total"] < 50:
        continue
    print(f"| `{k[:70]}` | {c['total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH[:17]) + " |")
