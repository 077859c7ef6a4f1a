#!/usr/bin/env python3
"""Hot-spot summary of `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > f.csv`: stall samples per
barrier-delimited segment and the top instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
iS, iW, iE = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = [(r[iS].strip(), int(r[iW] or 0), int(r[iE] or 0)) for r in rows[h + 1:] if len(r) > iE and r[0] != "Address"]
tot = sum(d[1] for d in data)
print("total samples", tot, "sass instructions", len(data), "executed", sum(d[2] for d in data))
bars = [i for i, d in enumerate(data) if d[0].startswith("BAR") or "WARPSYNC" in d[0]]
bounds = [0] + bars + [len(data)]
for a, b in zip(bounds, bounds[1:]):
    s = sum(d[1] for d in data[a:b])
    e = sum(d[2] for d in data[a:b])
    if s * 50 > tot:
        print("segment [%d,%d) %-28s samples %6d %5.1f%% executed %d" % (a, b, data[a][0][:28], s, 100 * s / max(tot, 1), e))
for i, d in sorted(enumerate(data), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%6d %6d %8d  %s" % (i, d[1], d[2], d[0][:100]))
