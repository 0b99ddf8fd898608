#!/bin/bash
# sparse-uniqueness timing of the product build and every build under tools/ab (hashsort shape sweeps)
for lib in term_b200/libtermgpu.so tools/ab/libtermgpu_*.so; do
  TG_LIB=$PWD/$lib python tools/bench_suites.py c4 --steps 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'sparse' in d['workload']: print('$lib', d['workload'], 'kernel_ms', round(d['kernel_ms'],3), d.get('metric'))
"
done
