#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (optionally only launches >= --skip)."""
import csv, sys, collections
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = open(path, newline="").read().splitlines()
hi = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[hi:]))
agg = collections.OrderedDict()
for r in rows:
    if int(r["ID"]) < skip:
        continue
    k = r["Kernel Name"].split("(")[0][-70:]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"])
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{a[0]:6d} x {a[1]/a[0]/1e3:10.1f} us = {a[1]/1e3:12.1f} us ({100*a[1]/tot:5.1f}%)  {k}")
print(f"total {tot/1e3:.1f} us")
