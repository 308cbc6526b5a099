#!/usr/bin/env python3
"""Joins the per-instruction counters of an ncu capture (`ncu -i X.ncu-rep --page source --csv --print-source sass`)
with the line table of the same kernel (`nvdisasm -g` of the cubin extracted with `cuobjdump -xelf all`) and
prints executed warp instructions and stall samples per source region.

    python tools/hot_lines.py <source.csv> <nvdisasm -g output> <mangled kernel name> [lines per bucket]
"""
import collections
import csv
import re
import sys

src_csv, dis, kern = sys.argv[1:4]
bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 100
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ie, iss = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
counters = [(int(r[ie]), int(r[iss])) for r in rows[2:] if len(r) > ie]

# line table: the .text section of the kernel, instructions in address order with "//## File ..., line N" before them
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.strip() == ".text." + kern + ":")
cur = ("?", 0)
per_instr = []
for l in lines[start + 1:]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+[A-Z@{]", l):
        per_instr.append(cur)
    if l.strip().startswith(".section") or l.startswith("//-----"):
        break
n = min(len(per_instr), len(counters))
print(f"{len(counters)} profiled instructions, {len(per_instr)} in the line table; joined {n}")
agg = collections.defaultdict(lambda: [0, 0, 0])
for (f, ln), (e, s) in zip(per_instr[:n], counters[:n]):
    k = (f, ln // bucket * bucket)
    agg[k][0] += e; agg[k][1] += s; agg[k][2] += 1
te = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"{'file':22s} {'lines':>11s} {'static':>7s} {'executed %':>10s} {'stall samples %':>15s}")
for (f, ln), (e, s, c) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{f:22s} {ln:5d}-{ln + bucket - 1:<5d} {c:7d} {e / te * 100:10.1f} {s / ts * 100:15.1f}")
