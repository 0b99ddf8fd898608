#!/bin/bash
# sweep of the partitioned hash job's bucket size / table load on C4 (is_unique + FK), one line per setting
mkdir -p gpurun_out
: > gpurun_out/hash_sweep.txt
for T in 1048576 2097152 4194304; do
  for F in 2 4; do
    echo "== bucket_keys=$T slots_factor=$F" >> gpurun_out/hash_sweep.txt
    TG_HASH_BUCKET_KEYS=$T TG_HASH_SLOTS_FACTOR=$F python tools/bench_suites.py c4 --steps 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['workload'], 'kernel_ms %.3f launches %d' % (d['kernel_ms'], d['launches']))" >> gpurun_out/hash_sweep.txt
  done
done
cat gpurun_out/hash_sweep.txt
