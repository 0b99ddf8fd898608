#!/bin/bash
# Spearman / KLL timing of every library build under tools/ab (+ the product build): radix-pass shape sweeps on one box
for lib in term_b200/libtermgpu.so tools/ab/libtermgpu_*.so; do
  TG_LIB=$PWD/$lib python tools/bench_suites.py sp --steps 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$lib', d['workload'], 'kernel_ms', round(d['kernel_ms'],3), d.get('rho'))
"
done
