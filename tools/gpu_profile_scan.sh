#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, headline bench, ncu launch list, ncu full capture of scan_kernel.
set -x
mkdir -p gpurun_out
TAG=${1:-r1d}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_$TAG.log
python bench.py --variants > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/launches_$TAG.log 2>&1
TG_BENCH_ROWS=20000000 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 \
    -o gpurun_out/scan_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_${TAG}.log 2>&1
TG_BENCH_ROWS=20000000 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 8 -c 1 \
    -o gpurun_out/scan_${TAG}_full -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --variants > gpurun_out/ncu_${TAG}_full.log 2>&1
tail -c 600 gpurun_out/pytest_$TAG.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"], d.get("variants"))
PY
