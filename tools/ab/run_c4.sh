#!/bin/bash
# C4 (dense + sparse uniqueness, foreign key) timing of the product build and every build under tools/ab
for lib in term_b200/libtermgpu.so tools/ab/libtermgpu_*.so; do
  TG_LIB=$PWD/$lib python tools/bench_suites.py c4 --steps 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$lib', d['workload'], 'kernel_ms', round(d['kernel_ms'],3), 'wall', round(d['wall_ms'],3), d.get('metric'))
"
done
