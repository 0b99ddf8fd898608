#!/bin/bash
# quick GPU check: parity tests + headline bench (+variants); optional extra command
mkdir -p gpurun_out
TAG=${1:-q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_$TAG.log
cat gpurun_out/pytest_$TAG.log
python bench.py --variants --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"])
print("e2e", d["e2e"]); print(d.get("variants"))
PY
