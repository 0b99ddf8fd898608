#!/bin/bash
# launch list + ncu --set full captures of the radix-sort kernels through the Spearman job (tools/bench_suites.py sp)
# usage: tools/gpu_profile_sort.sh TAG [scale]
TAG=${1:-x}
SCALE=${2:-1.0}
mkdir -p gpurun_out
python tools/bench_suites.py sp --scale $SCALE --steps 3 > gpurun_out/sp_${TAG}.jsonl 2> gpurun_out/sp_${TAG}.err
cat gpurun_out/sp_${TAG}.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sp_launches_${TAG}.csv \
    python tools/bench_suites.py sp --scale $SCALE --steps 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/sp_launches_${TAG}.csv 2>/dev/null | head -40
for pr in rs_pass_kernel:40 rs_hist_kernel:2; do
  IFS=: read k skip <<< "$pr"
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/${k}_${TAG} -f \
      python tools/bench_suites.py sp --scale $SCALE --steps 1 > gpurun_out/ncu_${k}_${TAG}.log 2>&1
done
ls -la gpurun_out/*${TAG}*.ncu-rep
