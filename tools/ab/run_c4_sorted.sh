#!/bin/bash
# sparse uniqueness: the sorted-bucket path (hashsort.cu) against the partitioned path it replaces as the first choice
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "uniqueness or distinct" 2>&1 | tail -4
for mode in sorted partitioned; do
  if [ $mode = partitioned ]; then export TG_HASH_NO_SORTED=1; else unset TG_HASH_NO_SORTED; fi
  python tools/bench_suites.py c4 --steps 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$mode', d['workload'], 'kernel_ms', round(d['kernel_ms'],3), 'wall', round(d['wall_ms'],3), d.get('metric'), 'launches', d.get('launches'))
"
done
