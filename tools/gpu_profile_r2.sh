#!/bin/bash
# Round-2 profile run on one GPU: launch list of the bench, fresh ncu --set full captures of the dominant kernels.
TAG=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --suites none > gpurun_out/launches_bench_$TAG.log 2>&1
TG_BENCH_ROWS=20000000 ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 3 -c 1 \
    -o gpurun_out/scan_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --suites none > gpurun_out/ncu_scan_${TAG}.log 2>&1
for pr in c5:group_count_dict_kernel:0 c5:kll_sample_kernel:2 c3:dfa_kernel:2 c4:dense_kernel:0; do
  IFS=: read w k skip <<< "$pr"
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/${k}_${TAG} -f \
      python tools/bench_suites.py $w --scale 0.4 --steps 1 > gpurun_out/ncu_${k}_${TAG}.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:rs_pass_kernel -s 20 -c 1 -o gpurun_out/rs_pass_kernel_${TAG} -f \
    python tools/bench_suites.py sp --scale 0.4 --steps 1 > gpurun_out/ncu_rs_pass_${TAG}.log 2>&1
ls -la gpurun_out/*${TAG}*.ncu-rep
