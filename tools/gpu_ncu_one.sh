#!/bin/bash
# one ncu --set full capture of kernel $2 (regex) from bench_suites workload $1 at scale $4 (default 0.4), skipping $3 launches
W=$1; K=$2; SKIP=${3:-0}; SCALE=${4:-0.4}; TAG=${5:-x}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o gpurun_out/${K}_${TAG} -f \
    python tools/bench_suites.py $W --scale $SCALE --steps 1 > gpurun_out/ncu_${K}_${TAG}.log 2>&1
ls -la gpurun_out/${K}_${TAG}.ncu-rep
