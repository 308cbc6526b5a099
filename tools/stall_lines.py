#!/usr/bin/env python3
"""stall samples of one kind per source line: joins an `ncu --page source --csv --print-source sass` export with the
line table of the same build (`nvdisasm -g` of the cubin from `cuobjdump -xelf all ngb_cuda.o`).

    python tools/stall_lines.py <source.csv> <nvdisasm -g output> <mangled kernel> [stall column] [top N]"""
import collections
import csv
import re
import sys

src_csv, dis, kern = sys.argv[1:4]
col = sys.argv[4] if len(sys.argv) > 4 else "stall_long_sb"
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ie, il, isrc = h.index("Instructions Executed"), h.index(col), h.index("Source")
R = [r for r in rows[2:] if len(r) > ie]
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.strip() == ".text." + kern + ":")
cur = ("?", 0)
per = []
for l in lines[start + 1:]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+[A-Z@{]", l):
        per.append(cur)
    if l.strip().startswith(".section") or l.startswith("//-----"):
        break
print(f"{len(R)} profiled instructions, {len(per)} in the line table")
agg = collections.Counter()
aggop = collections.defaultdict(collections.Counter)
for (f, ln), r in zip(per, R):
    n = int(r[il])
    if n:
        agg[(f, ln)] += n
        aggop[(f, ln)][re.sub(r"^\s*(@!?U?P\d+\s+)?", "", r[isrc]).split()[0]] += n
tot = sum(agg.values())
for k, v in agg.most_common(top):
    print(f"{k[0]:20s} {k[1]:5d} {v:6d} {v / tot * 100:5.1f}%  {dict(aggop[k])}")
