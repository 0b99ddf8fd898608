"""Raw concurrent host->device ceiling of the box: N processes (one per GPU, torchrun), each copies 3.25 GB (the C2 e2e
payload: 4 x (800 MB values + 12.5 MB validity)) from pinned host memory to its GPU with plain cudaMemcpyAsync, all
ranks at once. Prints per-rank and aggregate GB/s — the number the engine's e2e path is bounded by at N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29620 tools/micro/h2d_concurrent.py
"""
import json
import os
import time

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
sizes = [800_000_000, 12_500_000] * 4
host = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes]
for h in host:
    h.zero_()
devb = [torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes]
total = sum(sizes)
out = {}
for n_streams in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    best = 1e9
    for rep in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, (h, d) in enumerate(zip(host, devb)):
            with torch.cuda.stream(streams[i % n_streams]):
                d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    out[f"streams_{n_streams}"] = {"ms": best * 1e3, "per_rank_gbs": total / best / 1e9, "aggregate_gbs": total * world / best / 1e9}
if rank == 0:
    print(json.dumps({"check": "h2d_concurrent_ceiling", "world": world, "bytes_per_rank": total, "host_cpus": os.cpu_count(), **out}), flush=True)
if world > 1:
    dist.destroy_process_group()
