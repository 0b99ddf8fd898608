#!/bin/bash
# sparse-key (radix-partitioned) is_unique on C4 versus the number of concurrent bucket pipelines
for L in 1 2 4; do for F in 2 4; do echo "lanes=$L slots_factor=$F"; TG_HASH_NO_DENSE=1 TG_HASH_LANES=$L TG_HASH_SLOTS_FACTOR=$F python tools/bench_suites.py c4 --steps 3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', d['workload'], 'kernel_ms %.3f' % d['kernel_ms'], d.get('metric'))"; done; done
