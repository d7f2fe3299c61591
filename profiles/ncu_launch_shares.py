#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: total time,
share, launch count.  Durations under ncu are cold-cache and serialised: compare SHARES."""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    name = re.sub(r"\(.*", "", r[ik])[:90]
    agg[name][0] += v
    agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms:10.3f} ms {100 * ms / tot:5.1f}%  n={n:5d}  {name}")
print(f"total ms {tot:.3f}")
