#!/bin/bash
# grouped completeness only: parity subset (bounded) + timing of product build vs tools/ab builds, every step under its own timeout
timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_merge.py tests/test_histogram_constraint.py -m gpu -x -q -k "group or histogram" 2>&1 | tail -3
for lib in term_b200/libtermgpu.so tools/ab/libtermgpu_*.so; do
  TG_LIB=$PWD/$lib timeout 90 python tools/bench_suites.py c5 --steps 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'grouped' in str(d.get('workload')): print('$lib', d.get('workload'), 'kernel_ms', round(d.get('kernel_ms',0) or 0,3), 'wall', round(d.get('wall_ms',0) or 0,3))
"
done
