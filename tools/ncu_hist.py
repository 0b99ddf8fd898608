#!/usr/bin/env python3
"""Histogram of warp-instructions by per-instruction execution count (identifies inner loops) for an .ncu-rep"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isrc, ist, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[ist] or 0), int(r[iex] or 0), r[ia][-5:], r[isrc]))
    except Exception:
        pass
tot = sum(d[1] for d in data)
print("total warp-instr", tot)
# contiguous regions with exec count > threshold
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.002
reg = []
cur = None
for d in data:
    if d[1] >= thr * tot:
        if cur is None: cur = [d[2], d[2], 0, 0, 0, collections.Counter()]
        cur[1] = d[2]; cur[2] += d[1]; cur[3] += 1; cur[4] += d[0]; cur[5][d[3].split()[0] if not d[3].startswith('@') else d[3].split()[1]] += 1
    else:
        if cur: reg.append(cur); cur = None
if cur: reg.append(cur)
for r in reg:
    print(f"{r[0]}-{r[1]}  ninstr={r[3]:4d} exec={r[2]:>11} ({100*r[2]/tot:5.1f}%) samples={r[4]:6d}  ", dict(r[5].most_common(8)))
