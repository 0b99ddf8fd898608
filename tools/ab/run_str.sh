#!/bin/bash
# string kernel: byte-indexed (direct) tables against class-indexed tables on C1 / C3 / the mixed suite
for mode in direct class; do
  if [ $mode = class ]; then export TG_STR_NO_DIRECT=1; else unset TG_STR_NO_DIRECT; fi
  python tools/bench_suites.py c1 c3 x --steps 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$mode', d.get('workload'), 'kernel_ms', round(d.get('kernel_ms',0) or 0,3), 'wall', round(d.get('wall_ms',0) or 0,3), {k: round(v,3) for k,v in d.items() if k.endswith('_ms') and isinstance(v,float)})
"
done
