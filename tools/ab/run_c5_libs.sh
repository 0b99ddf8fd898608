#!/bin/bash
# C5 (KLL / grouped completeness / Spearman) timing of the product build and every build under tools/ab
for lib in term_b200/libtermgpu.so tools/ab/libtermgpu_*.so; do
  TG_LIB=$PWD/$lib python tools/bench_suites.py c5 --steps 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$lib', d.get('workload'), 'kernel_ms', round(d.get('kernel_ms',0) or 0,3), 'wall', round(d.get('wall_ms',0) or 0,3))
"
done
