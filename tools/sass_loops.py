#!/usr/bin/env python3
"""List the innermost loops (backward branches) of a SASS dump with their instruction mix."""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
ins = []
for ln in lines:
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2i = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.\w+)* (?:`\(\.L_x_\d+\)|0x([0-9a-f]+))', t)
    if m and m.group(1):
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2i:
            loops.append((addr2i[tgt], i))
# innermost only
inner = [l for l in loops if not any((o[0] >= l[0] and o[1] <= l[1] and o != l) for o in loops)]
want = sys.argv[2] if len(sys.argv) > 2 else None
for s, e in inner:
    body = [t for _, t in ins[s:e + 1]]
    ops = collections.Counter((t.split()[1] if t.startswith('@') else t.split()[0]) for t in body)
    if want and not any(want in t for t in body):
        continue
    print(f"{ins[s][0]:06x}-{ins[e][0]:06x} n={len(body):4d}", dict(ops.most_common(14)))
