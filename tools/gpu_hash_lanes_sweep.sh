#!/bin/bash
# sparse-key (radix-partitioned) is_unique on C4: concurrent bucket pipelines x bucket size x table slots per key
for L in 2 3 4; do for T in 1048576 2097152 4194304; do for F in 4 8; do echo -n "lanes=$L bucket_keys=$T slots_factor=$F: "; TG_HASH_NO_DENSE=1 TG_HASH_LANES=$L TG_HASH_BUCKET_KEYS=$T TG_HASH_SLOTS_FACTOR=$F python tools/bench_suites.py c4 --steps 3 2>/dev/null | python -c "
import sys, json
print(' | '.join('%s %.3f ms' % (json.loads(l)['workload'], json.loads(l)['kernel_ms']) for l in sys.stdin))"; done; done; done
