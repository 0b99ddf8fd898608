#!/bin/bash
# full single-GPU check: every GPU test, then the bench line (with suites) and the reference arm
TAG=${1:-f}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 240 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["ms_per_step"], d.get("parity_vs_cpu_arm"))
for s in d.get("suites", []):
    print(s["name"], "ms", round(s.get("ms_per_step", 0), 3), "kms", round(s.get("kernel_ms_rank0", 0) or 0, 3), "frac", round(s.get("frac", 0), 4), s.get("error", ""))
PY
