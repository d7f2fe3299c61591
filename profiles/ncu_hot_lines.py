#!/usr/bin/env python
"""Top warp-stall SASS lines from `ncu -i X.ncu-rep --page source --csv --kernel-name ...`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
hdr = rows[hi]
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
data = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) <= i_ex or not r[i_s].isdigit():
        continue
    data.append((int(r[i_s]), int(r[i_ex] or 0), r[i_src].strip(), n))
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "sass lines", len(data), "instr executed", sum(d[1] for d in data))
for s, ex, src, n in sorted(data, key=lambda d: -d[0])[:top]:
    print(f"{s:7d} {100 * s / tot:5.1f}%  ex={ex:10d}  #{n:5d} {src[:100]}")
