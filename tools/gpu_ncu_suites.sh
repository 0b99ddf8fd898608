#!/bin/bash
# ncu --set full captures of the secondary kernels (one launch each) through tools/bench_suites.py at reduced scale
# usage: tools/gpu_ncu_suites.sh TAG "c3:dfa_kernel[:skip] c4:insert64_kernel ..." [scale]
TAG=${1:-x}
PAIRS=${2:-"c3:dfa_kernel"}
SCALE=${3:-0.4}
mkdir -p gpurun_out
for pr in $PAIRS; do
  IFS=: read w k skip <<< "$pr"
  skip=${skip:-2}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/${k}_${TAG}_s${skip} -f \
      python tools/bench_suites.py $w --scale $SCALE --steps 1 > gpurun_out/ncu_${k}_${TAG}_s${skip}.log 2>&1
done
ls -la gpurun_out/*${TAG}*.ncu-rep
