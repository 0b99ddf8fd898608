#!/bin/bash
# environment sweep of the sparse-key hash path (no rebuild): lanes x bucket keys x slots factor
for lanes in 2 4; do for bk in 2097152 4194304 8388608; do for sf in 2 4; do
  TG_HASH_LANES=$lanes TG_HASH_BUCKET_KEYS=$bk TG_HASH_SLOTS_FACTOR=$sf python tools/bench_suites.py c4 --steps 3 2>/dev/null | python -c "
import sys,json
out=[]
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'sparse' in d['workload']: out.append((round(d['kernel_ms'],3), d.get('launches')))
print('lanes=$lanes bucket_keys=$bk slots_factor=$sf sparse', out)
"
done; done; done
