#!/usr/bin/env python3
"""Executed-instruction mix of one kernel from `ncu --page source --csv` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
seen = set()
mix = collections.Counter()
stall = collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= iE or r[0] == "Address" or r[0] in seen:
        continue
    seen.add(r[0])
    op = r[iS].strip().split()
    if op and op[0].startswith("@"):
        op = op[1:]
    name = op[0].split(".")[0] if op else "?"
    mix[name] += int(r[iE] or 0)
    stall[name] += int(r[iW] or 0)
tot = sum(mix.values())
ts = sum(stall.values())
print("executed warp-instructions", tot, "stall samples", ts)
for k, v in mix.most_common(28):
    print("%-10s %10d %5.1f%%   stall %5.1f%%" % (k, v, 100 * v / tot, 100 * stall[k] / max(ts, 1)))
