#!/usr/bin/env python3
"""Aggregate an .ncu-rep's source page by CUDA source line: executed warp-instructions and stall samples per line."""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iex, ist = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg, cur = {}, None
for r in rows[hi + 1:]:
    if not r:
        continue
    if r[0] not in ("", "File Path", "Function Name", "Line No"):
        cur = (r[0], r[1].strip()[:110])
        continue
    if r[0] == "" and cur is not None and len(r) > max(iex, ist):
        try:
            a = agg.setdefault(cur, [0, 0])
            a[0] += int(r[iex] or 0)
            a[1] += int(r[ist] or 0)
        except ValueError:
            pass
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100 * a[0] / tot_i:5.1f}% instr {100 * a[1] / tot_s:5.1f}% samples  L{ln}: {src}")
