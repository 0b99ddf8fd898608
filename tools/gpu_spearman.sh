#!/bin/bash
TAG=${1:-s}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "spearman or rank or dist" 2>&1 | tail -8
python tools/bench_suites.py sp --steps 3 2>gpurun_out/sp_$TAG.err | tee gpurun_out/sp_$TAG.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['workload'], 'kernel_ms', round(d['kernel_ms'],3), 'wall', round(d['wall_ms'],3), 'frac', round(d['frac'],4), d.get('rho'))
"
tail -3 gpurun_out/sp_$TAG.err
