#!/bin/bash
# ncu full capture of scan_kernel: literal suite (TAG) and full numeric set (TAG_full), 20M rows
TAG=${1:-x}
mkdir -p gpurun_out
TG_BENCH_ROWS=20000000 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 \
    -o gpurun_out/scan_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_${TAG}.log 2>&1
if [ -z "$2" ]; then
TG_BENCH_ROWS=20000000 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 8 -c 1 \
    -o gpurun_out/scan_${TAG}_full -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --variants > gpurun_out/ncu_${TAG}_full.log 2>&1
fi
