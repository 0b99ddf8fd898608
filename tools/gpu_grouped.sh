#!/bin/bash
# grouped-completeness iteration: parity tests that touch the grouped path, C5 lines, launch list
TAG=${1:-g}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "group or merge or mixed" 2>&1 | tail -5
python tools/bench_suites.py c5 --steps 5 2>gpurun_out/c5_$TAG.err | tee gpurun_out/c5_$TAG.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['workload'], d.get('config'), 'kernel_ms', round(d['kernel_ms'],3), 'wall', round(d['wall_ms'],3), 'frac', round(d['frac'],4))
"
tail -3 gpurun_out/c5_$TAG.err
