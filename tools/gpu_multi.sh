#!/bin/bash
# N-GPU check on one box: world-N parity (tools/dist_check.py) + the bench line with every suite
N=${1:-2}; TAG=${2:-m}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/dist_n${N}_$TAG.log 2>&1
tail -4 gpurun_out/dist_n${N}_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 100 --warmup 3 \
    > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
tail -2 gpurun_out/bench_n${N}_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n${N}_$TAG.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
for s in d.get("suites", []):
    c = s.get("collectives", {})
    print(s["name"], "ms", round(s.get("ms_per_step", 0), 3), "kms", round(s.get("kernel_ms_rank0", 0) or 0, 3), "frac", round(s.get("frac", 0), 4),
          "shuffle_ms", round(c.get("shuffle_ms_per_step_max_rank", 0), 3), "exch_ms", round(c.get("partial_exchange_ms_per_step_max_rank", 0), 3), s.get("error", ""))
PY
