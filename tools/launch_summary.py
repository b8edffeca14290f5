#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches and mean / total time."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    a = agg.setdefault(r[ki][:110], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
total = sum(t for _, t in agg.values())
print(f"total {total / 1e6:.3f} ms over {sum(c for c, _ in agg.values())} launches")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t / c / 1000:10.1f} us x{c:5d} {100 * t / total:5.1f}%  {n}")
