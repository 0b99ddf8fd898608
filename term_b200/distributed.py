"""Row-partitioned multi-GPU execution: one process per GPU, each rank evaluates the
plan on its own row shard, the serialised partial aggregates are all-gathered (torch.distributed: NCCL
on GPUs, gloo in the CPU tests) and merged IN RANK ORDER on every rank, so all ranks finalize to the
same, deterministic result. Mirrors AnalyzerState::merge (analyzers/traits.rs:154-179) and the
IncrementalAnalysisRunner's partition -> state -> merge flow (analyzers/incremental/runner.rs:165-358).
"""
import torch
import torch.distributed as dist


def allgather_blobs(blob: bytes, device=None):
    """All-gather variable-length byte strings; returns the list ordered by rank."""
    world = dist.get_world_size()
    backend = dist.get_backend()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))
    n = torch.tensor([len(blob)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes)
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if len(blob):
        buf[: len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    out = [torch.zeros(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(out, buf)
    return [bytes(o[:s].cpu().numpy().tobytes()) for o, s in zip(out, sizes)]


def merge_partials(plan, blobs):
    """Reset the plan's partial state and merge every rank's blob in rank order, then finalize."""
    plan.partial_reset()
    for b in blobs:
        plan.partial_merge(b)
    plan.finalize()


def execute_distributed(plan, ctx, table="data"):
    """Each rank: partial execute on its shard -> all-gather -> ordered merge -> finalize."""
    plan.execute_partial(ctx, table)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        plan.finalize()
        return
    blobs = allgather_blobs(plan.partial_export())
    merge_partials(plan, blobs)
