#!/bin/bash
# Round-2 (second half) profile run on one GPU: launch list of the bench (headline + suites), ncu --set full of the kernels that changed
TAG=${1:-r2b}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --suites c3,c4,c5,mixed > gpurun_out/launches_bench_$TAG.log 2>&1
for pr in c5:group_count_tile_kernel:0:c5 sp:rs_pass_kernel:6:sp sp:rk_rank_x_kernel:0:sp; do
  IFS=: read w k skip name <<< "$pr"
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/${k}_${TAG} -f \
      python tools/bench_suites.py $w --scale 0.4 --steps 1 > gpurun_out/ncu_${k}_${TAG}.log 2>&1
done
ls -la gpurun_out/*${TAG}*.ncu-rep
