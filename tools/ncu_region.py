#!/usr/bin/env python3
"""Print SASS with samples / exec counts for an address range of an .ncu-rep: ncu_region.py rep 0x14800 0x15520"""
import csv, subprocess, sys, io
rep, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, isrc, ist, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
for r in rows[h + 1:]:
    try:
        a = int(r[ia], 16)
    except Exception:
        continue
    if base is None:
        base = a
    off = a - base
    if lo <= off <= hi:
        print(f"{off:06x} smp={int(r[ist] or 0):5d} exec={int(r[iex] or 0):8d}  {r[isrc][:90]}")
