#!/usr/bin/env python3
"""Breaks one distributed step of the headline suite into its parts (run under torchrun): where do the ~0.1 ms
between the 1-GPU and the N-GPU step go?"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench as B
import term_b200 as T
from term_b200 import distributed as D

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = T.SessionContext(local)
n = 100_000_000
cols, keep = B.make_device_table(torch, n, 42 + rank, dev)
ctx.register_device_table("data", {k: {kk: vv for kk, vv in v.items() if kk not in ("tensor", "bits")} for k, v in cols.items()}, keepalive=keep)
plan, slots = B.build_suite(T, "data").build_plan()
for _ in range(5):
    D.execute_distributed(plan, ctx, "data")
dist.barrier(); torch.cuda.synchronize()
acc = {"execute_partial": 0.0, "export": 0.0, "exchange": 0.0, "merge": 0.0, "total": 0.0}
K = 200
for _ in range(K):
    t0 = time.perf_counter()
    plan.execute_partial(ctx, "data")
    t1 = time.perf_counter()
    blob = plan.partial_export(); aggs = plan.aggregates()
    t2 = time.perf_counter()
    cap = (8 + len(aggs) * 160 + 1024 + 4095) // 4096 * 4096
    got = D.allgather_blobs_fixed(blob, cap)
    t3 = time.perf_counter()
    D.merge_partials(plan, got)
    t4 = time.perf_counter()
    for k, v in zip(("execute_partial", "export", "exchange", "merge", "total"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)):
        acc[k] += v
if rank == 0:
    print(json.dumps({k: round(v / K * 1e6, 1) for k, v in acc.items()}), "us per step, world", world, flush=True)
ctx.close(); dist.destroy_process_group()
