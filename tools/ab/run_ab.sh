#!/bin/bash
# A/B of two builds of the library on the headline bench (kernel-parameter / code-change checks on one box)
for rep in 1 2; do
for lib in term_b200/libtermgpu.so tools/ab/libtermgpu_prevscan.so; do
  TG_LIB=$PWD/$lib python bench.py --no-cpu --no-e2e --suites c2full --steps 200 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), [ (s['name'], round(s['ms_per_step'],4)) for s in d.get('suites',[])])
"
done
done
for lib in term_b200/libtermgpu.so tools/ab/libtermgpu_prevscan.so; do TG_LIB=$PWD/$lib python tools/ab/pred_time.py 2>&1 | tail -3; done
