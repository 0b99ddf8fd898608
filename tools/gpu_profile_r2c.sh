#!/bin/bash
# Round-2 (final) profile run on one GPU: launch list of the bench (headline + every suite), ncu --set full of the kernels that
# changed since r2b (the rank kernels, the scan kernel after the evaluator change)
TAG=${1:-r2c}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --suites c1,c2full,c3,c4,c5,mixed > gpurun_out/launches_bench_$TAG.log 2>&1
for pr in sp:rk_rank_x_kernel:0:sp sp:rk_rank_y_moments_kernel:0:sp; do
  IFS=: read w k skip name <<< "$pr"
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/${k}_${TAG} -f \
      python tools/bench_suites.py $w --scale 0.4 --steps 1 > gpurun_out/ncu_${k}_${TAG}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 4 -c 1 -o gpurun_out/scan_kernel_${TAG} -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --suites none > gpurun_out/ncu_scan_kernel_${TAG}.log 2>&1
ls -la gpurun_out/*${TAG}*.ncu-rep
