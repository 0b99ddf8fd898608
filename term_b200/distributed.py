"""Multi-GPU execution: one process per GPU, every rank holds a row shard of each table.

* Scan / string / KLL / grouped aggregates shard by rows: each rank evaluates the plan on its shard, the
  serialised partial aggregates are all-gathered (torch.distributed: NCCL on GPUs, gloo in the CPU tests) and
  merged IN RANK ORDER on every rank, so all ranks finalize to the same, deterministic result. Mirrors
  AnalyzerState::merge (analyzers/traits.rs:154-179) and the IncrementalAnalysisRunner's partition -> state ->
  merge flow (analyzers/incremental/runner.rs:165-358).
* COUNT(DISTINCT ..) / uniqueness and the foreign-key anti-join do not merge by rows. Their keys are first hash-
  shuffled — the step DataFusion's RepartitionExec(Hash) performs under the reference's SQL
  (constraints/uniqueness.rs:549-718, foreign_key.rs:165-172): every rank groups its keys by destination rank on
  the device (tg_table_partition_keys), one all-to-all moves them (NCCL over NVLink), the receiving rank adopts
  its keys as a table and the aggregate is redirected to it (tg_plan_redirect_aggregate). Equal keys now live on
  exactly one rank, so the per-rank states add up exactly. NULL rows travel as a count to rank 0. Utf8 and composite
  keys travel as their 128-bit fingerprints (tg_table_partition_fingerprints), the identity the single-GPU path uses.
* Spearman needs GLOBAL ranks (RANK() OVER (ORDER BY ..) over the whole table, analyzers/advanced/correlation.rs:
  334-350), so per-shard rank sums do not add up. It runs as a distributed SAMPLE SORT, once per
  column: every rank sorts its shard's pairwise-complete rows (tg_rank_local_sort), contributes evenly spaced sample
  keys, all ranks derive the same world - 1 splitters, one all-to-all by key range (NCCL over NVLink) brings every
  range to its owner, which sorts it and turns local run heads into global minimum ranks by adding the number of keys
  on the lower ranks (tg_rank_finish_x / _y). The ranks of x travel with the rows through the second sort; each rank
  ends with the shifted rank co-moments of its y range, which merge by addition. Equal keys always meet on one rank,
  so ties get the same global minimum rank as on one GPU. (The round-1 path — gather every pair to rank 0 — is kept
  as gather_pairs / TG_SPEARMAN_GATHER=1 for comparison.)
"""
import os

import torch
import torch.distributed as dist

from . import _ffi as F

KIND_DISTINCT, KIND_FK, KIND_SPEARMAN = 6, 7, 10

# bench.py sets this to a dict to collect, per timed region, the bytes every rank sends through the NVLink all-to-all
# (shuffle_bytes), the device time of those collectives (shuffle_ms, CUDA events) and the wall time of the partial-state
# exchange (exchange_ms). None: no bookkeeping at all.
PROFILE = None


def _device(device=None):
    if device is not None:
        return device
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def bind_to_gpu_numa(device_index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `device_index`, so that the pinned host buffers it
    allocates afterwards (first touch) and its copy-issuing threads sit on the GPU's own socket: with one process per
    GPU all ranks stage host->device concurrently and a remote-socket buffer halves a rank's H2D rate. Call it before
    the SessionContext and any pinned allocation. Returns the CPU set bound, or None when nothing was changed (NVML
    unavailable, TG_NUMA_BIND=0, the container's cpuset does not intersect the GPU's, or no narrowing)."""
    if os.environ.get("TG_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        local = {i for i in range(n_cpus) if (int(words[i // 64]) >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        if not target or target == allowed:
            return None
        os.sched_setaffinity(0, target)
        return sorted(target)
    except Exception:  # affinity is an optimisation: never fail the run over it
        return None


def allgather_blobs(blob: bytes, device=None):
    """All-gather variable-length byte strings; returns the list ordered by rank."""
    world = dist.get_world_size()
    dev = _device(device)
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


_FIXED_KINDS = {0, 1, 2, 3, 4, 5, 6, 10, 11}  # aggregates whose partial record has a fixed size (no blob)
_SCAN_KINDS = {0, 1, 2, 3, 4}  # ROWS / VALID / NUM / PAIR / PRED: what the fused numeric scan answers (tg_plan_execute_exchange)
_bufs = {}


def allgather_blobs_fixed(blob: bytes, cap: int, device=None):
    """One-collective all-gather for blobs whose size every rank can bound WITHOUT talking (a plan of fixed-size
    aggregates): each rank sends [u64 length | payload] padded to `cap` bytes through cached pinned / device
    buffers. Returns None when some rank's blob did not fit (all ranks see that and take the general path)."""
    world = dist.get_world_size()
    dev = _device(device)
    key = (cap, str(dev), world)
    b = _bufs.get(key)
    if b is None:
        pin = dev.type == "cuda"
        b = dict(h_send=torch.zeros(cap, dtype=torch.uint8, pin_memory=pin), h_recv=torch.zeros(cap * world, dtype=torch.uint8, pin_memory=pin),
                 d_send=torch.zeros(cap, dtype=torch.uint8, device=dev), d_recv=torch.zeros(cap * world, dtype=torch.uint8, device=dev))
        _bufs[key] = b
    n = len(blob)
    fits = n + 8 <= cap
    hs = b["h_send"].numpy()
    hs[:8] = memoryview((n if fits else 0xFFFFFFFFFFFFFFFF).to_bytes(8, "little"))
    if fits:
        hs[8: 8 + n] = memoryview(blob)
    if dev.type == "cuda":
        b["d_send"].copy_(b["h_send"], non_blocking=True)
        dist.all_gather_into_tensor(b["d_recv"], b["d_send"])
        b["h_recv"].copy_(b["d_recv"], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        raw = b["h_recv"].numpy()
    else:
        outs = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(outs, b["h_send"])
        raw = torch.cat(outs).numpy()
    out = []
    for r in range(world):
        ln = int.from_bytes(raw[r * cap: r * cap + 8].tobytes(), "little")
        if ln == 0xFFFFFFFFFFFFFFFF:
            return None
        out.append(raw[r * cap + 8: r * cap + 8 + ln].tobytes())
    return out


MAILBOX_SLOT_BYTES = 1 << 20  # 1 MiB per rank and parity: KLL sketches / grouped tables fit; the collect kernel only moves what a slot holds


def _ensure_mailbox(ctx) -> bool:
    """One-time setup of the NVLink peer mailboxes (tg_engine_mailbox_*): create, all-gather the 64-byte CUDA IPC
    handles, open. Returns False (and remembers it) when the platform refuses, so the NCCL all-gather path is used."""
    import ctypes as C
    state = getattr(ctx, "_mailbox_state", None)
    world, rank = dist.get_world_size(), dist.get_rank()
    if state is not None:
        return state == ("ok", world)
    ok = True
    handle = C.create_string_buffer(64)
    try:
        if os.environ.get("TG_NO_MAILBOX"):
            raise RuntimeError("disabled")
        F.check(F.lib().tg_engine_mailbox_create(ctx.handle, world, rank, MAILBOX_SLOT_BYTES, handle))
    except Exception:
        ok = False
    dev = _device()
    mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(dev)
    allh = torch.zeros(64 * world, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh, mine)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item())
    if ok:
        try:
            raw = bytes(allh.cpu().numpy().tobytes())
            F.check(F.lib().tg_engine_mailbox_open(ctx.handle, raw))
        except Exception:
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    ctx._mailbox_state = ("ok", world) if ok else ("failed", world)
    return ok


def exchange_and_finalize(plan, ctx):
    """Exchange this rank's partial states with every rank, merge in rank order, finalize. Over NCCL on one node the
    blobs travel through the peer mailboxes (two tiny kernels over NVLink peer memory, see term_b200/csrc/mailbox.cu),
    whatever their size: a rank whose blob does not fit its slot publishes a marker instead, EVERY rank sees it and all
    take the NCCL all-gather together. Platforms without CUDA IPC (and the gloo tests) always take the all-gather."""
    import ctypes as C
    if dist.get_backend() == "nccl" and _ensure_mailbox(ctx):
        fell_back = C.c_int32(0)
        F.check(F.lib().tg_plan_exchange_ex(ctx.handle, plan.handle, C.byref(fell_back)))
        if not fell_back.value:
            return
    merge_partials(plan, exchange_partials(plan))


def execute_exchange_fused(plan, ctx, table):
    """tg_plan_execute_exchange: the whole step of a scan-only plan in one call (partial states assembled on the device,
    published over NVLink, one synchronisation). Returns False when the plan does not qualify (nothing was executed)."""
    import ctypes as C
    if dist.get_backend() != "nccl" or os.environ.get("TG_NO_FUSED_EXCHANGE") or not _ensure_mailbox(ctx):
        return False
    done = C.c_int32(0)
    F.check(F.lib().tg_plan_execute_exchange(ctx.handle, plan.handle, table.encode(), C.byref(done)))
    return bool(done.value)


def exchange_partials(plan):
    """All-gather this rank's partial blob; one collective when the plan's partial records are fixed-size."""
    blob = plan.partial_export()
    aggs = plan.aggregates()
    if all(k in _FIXED_KINDS for k, _ in aggs):
        # 8 (count) + per aggregate 8 + 8 + 64 + 64 + 8 + 8, plus room for an error message or two
        cap = (8 + len(aggs) * 160 + 1024 + 4095) // 4096 * 4096
        got = allgather_blobs_fixed(blob, cap)
        if got is not None:
            return got
    return allgather_blobs(blob)


def shuffle_keys(keys: torch.Tensor, counts, n_nulls: int):
    """All-to-all of hash-partitioned keys. `keys` (int64, grouped by destination rank) holds counts[r] keys for
    rank r. Returns (the keys this rank owns after the exchange, NULL rows this rank accounts for): every rank's
    NULL count goes to rank 0. Works on NCCL (device tensors) and gloo (CPU tensors)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = keys.device
    prof = PROFILE if (PROFILE is not None and dev.type == "cuda") else None
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    send = torch.tensor(list(counts), dtype=torch.int64, device=dev)
    recv = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send)
    recv_counts = [int(x) for x in recv.tolist()]
    out = torch.empty(sum(recv_counts), dtype=torch.int64, device=dev)
    dist.all_to_all_single(out, keys[: sum(counts)].contiguous(), output_split_sizes=recv_counts, input_split_sizes=list(counts))
    nulls = torch.tensor([int(n_nulls)], dtype=torch.int64, device=dev)
    dist.all_reduce(nulls)
    if prof is not None:
        ev1.record()
        ev1.synchronize()
        prof["shuffle_ms"] += ev0.elapsed_time(ev1)
        prof["shuffle_bytes"] += 8 * (sum(counts) - counts[rank])  # what leaves this GPU
    return out, (int(nulls.item()) if rank == 0 else 0)


def _adopt_shard(ctx, name, column, dtype, keys: torch.Tensor, n_nulls: int):
    """Register this rank's shuffled keys (+ n_nulls NULL rows at the end) as a one-column device table."""
    n = keys.numel() + n_nulls
    vals = torch.zeros(n + 64, dtype=torch.int64, device=keys.device)
    vals[: keys.numel()] = keys
    keep = [vals]
    validity = None
    if n_nulls:
        bits = torch.zeros((n + 7) // 8 + 64, dtype=torch.uint8, device=keys.device)
        full, rem = divmod(keys.numel(), 8)
        bits[:full] = 0xFF
        if rem:
            bits[full] = (1 << rem) - 1
        keep.append(bits)
        validity = bits.data_ptr()
    ctx.register_device_table(name, {column: dict(dtype=dtype, n_rows=n, values=vals.data_ptr(), validity=validity)}, keepalive=keep)


def _column_dtype(ctx, table, column):
    return ctx.column_dtype(table, column)


def _ensure_comm(ctx) -> bool:
    """One-time setup of the library's own NCCL communicator (tg_comm_*): rank 0 draws the ncclUniqueId, torch.distributed
    carries its 128 bytes, every rank joins. Returns False (and remembers it) when the platform refuses or
    TG_NO_CABI_SHUFFLE is set: the shuffle then runs through torch.distributed's all_to_all_single."""
    import ctypes as C
    state = getattr(ctx, "_comm_state", None)
    world, rank = dist.get_world_size(), dist.get_rank()
    if state is not None:
        return state == ("ok", world)
    ok = dist.get_backend() == "nccl" and not os.environ.get("TG_NO_CABI_SHUFFLE")
    dev = _device()
    buf = C.create_string_buffer(128)
    if ok and rank == 0:
        try:
            F.check(F.lib().tg_comm_unique_id(buf))
        except Exception:
            ok = False
    idt = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(dev)
    dist.broadcast(idt, 0)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item())
    if ok:
        try:
            F.check(F.lib().tg_comm_init(ctx.handle, bytes(idt.cpu().numpy().tobytes()), world, rank))
        except Exception:
            ok = False  # a rank that fails here would leave the others inside ncclCommInitRank: NCCL reports it there
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    ctx._comm_state = ("ok", world) if ok else ("failed", world)
    return ok


def _cabi_shuffle(ctx, call, *args):
    """one tg_table_shuffle_* call with the PROFILE bookkeeping around it"""
    import ctypes as C
    import time
    n = C.c_int64()
    sent0 = F.lib().tg_comm_bytes_sent(ctx.handle)
    t0 = time.perf_counter()
    F.check(call(ctx.handle, *args, C.byref(n)))
    if PROFILE is not None:
        PROFILE["shuffle_ms"] += (time.perf_counter() - t0) * 1e3  # partition + counts + NCCL group + sync, one call
        PROFILE["shuffle_bytes"] += F.lib().tg_comm_bytes_sent(ctx.handle) - sent0
    return n.value


def _exact_shuffle_key(ctx, dtype):
    """the column travels as its own 64-bit values: Int64 / Float64 always; Int32 / Float32 when the library does the shuffle
    (tg_table_shuffle_column reads them through the exactly widened shadow) — otherwise as 128-bit fingerprints"""
    if dtype in (F.TG_INT64, F.TG_FLOAT64):
        return True
    return dtype in (F.TG_INT32, F.TG_FLOAT32) and _ensure_comm(ctx)


def _shuffle_column(ctx, table, column, shard_name, by_range_if_dense=False):
    """partition -> all-to-all -> adopt as table `shard_name` (column keeps its name). by_range_if_dense: a DISTINCT
    aggregate may split dense Int64 keys by value range (a foreign key may not: both sides must agree on the function)."""
    if _ensure_comm(ctx):  # NCCL inside the library: one call, parts land in the shard's own column buffer
        _cabi_shuffle(ctx, F.lib().tg_table_shuffle_column, table.encode(), column.encode(), shard_name.encode(), 1 if by_range_if_dense else 0)
        return
    world = dist.get_world_size()
    dtype = _column_dtype(ctx, table, column)
    if dtype not in (F.TG_INT64, F.TG_FLOAT64):
        raise NotImplementedError(f"multi-GPU shuffle of column '{table}.{column}': only Int64 / Float64 keys are supported")
    ptr, counts, nulls = ctx.partition_keys(table, column, world)
    total = sum(counts)
    dev = torch.device("cuda", torch.cuda.current_device())
    # view the engine-owned device buffer as a tensor (no copy); it is consumed before the next partition call
    if total:
        keys = _tensor_from_ptr(ptr, total, dev)
    else:
        keys = torch.empty(0, dtype=torch.int64, device=dev)
    mine, my_nulls = shuffle_keys(keys, counts, nulls)
    _adopt_shard(ctx, shard_name, column, dtype, mine, my_nulls)


def _shuffle_fingerprints(ctx, table, columns, shard_name):
    """Utf8 / composite keys: every row travels as its 24-byte fingerprint record {h1, h2, has_null}; the shard is
    adopted as one TG_FP128 column named tg_fp."""
    if _ensure_comm(ctx):
        import ctypes as C
        arr = (C.c_char_p * len(columns))(*[c.encode() for c in columns])
        _cabi_shuffle(ctx, F.lib().tg_table_shuffle_fingerprints, table.encode(), arr, len(columns), shard_name.encode())
        return
    world = dist.get_world_size()
    ptr, counts = ctx.partition_fingerprints(table, columns, world)
    total = sum(counts)
    dev = torch.device("cuda", torch.cuda.current_device())
    recs = _tensor_from_ptr(ptr, total * 3, dev) if total else torch.empty(0, dtype=torch.int64, device=dev)
    mine, _ = shuffle_keys(recs, [c * 3 for c in counts], 0)
    _adopt_fp_shard(ctx, shard_name, mine)


def complete_pairs(x: torch.Tensor, x_valid, y: torch.Tensor, y_valid) -> torch.Tensor:
    """Rows where both values are non-NULL, as an (m, 2) float64 tensor (CAST(.. AS DOUBLE), row order kept).
    x_valid / y_valid: bool masks or None (no NULLs)."""
    xd, yd = x.to(torch.float64), y.to(torch.float64)
    if x_valid is not None or y_valid is not None:
        m = torch.ones(x.numel(), dtype=torch.bool, device=x.device)
        if x_valid is not None:
            m &= x_valid
        if y_valid is not None:
            m &= y_valid
        xd, yd = xd[m], yd[m]
    return torch.stack([xd, yd], dim=1).contiguous()


def gather_pairs(pairs: torch.Tensor, dst: int = 0):
    """Concatenate every rank's (m_r, 2) float64 pairs on rank `dst`, in rank order. Other ranks get an empty tensor.
    One all-gather of the counts, then one padded gather (NCCL and gloo)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = pairs.device
    cnt = torch.tensor([pairs.shape[0]], dtype=torch.int64, device=dev)
    cnts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c.item()) for c in cnts]
    cap = max(max(cnts), 1)
    send = torch.zeros((cap, 2), dtype=torch.float64, device=dev)
    send[: pairs.shape[0]] = pairs
    recv = [torch.empty((cap, 2), dtype=torch.float64, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return torch.empty((0, 2), dtype=torch.float64, device=dev)
    return torch.cat([r[:c] for r, c in zip(recv, cnts)], dim=0)


def _unpack_validity(ptr, n, dev):
    """bool mask of the first n rows of an LSB-first validity bitmap at device address ptr"""
    bits = _tensor_from_ptr(ptr, (n + 7) // 8, dev, "|u1")
    sh = torch.arange(8, device=dev, dtype=torch.uint8)
    return ((bits.unsqueeze(1) >> sh) & 1).bool().view(-1)[:n]


def _gather_spearman_columns(ctx, table, cx, cy, name):
    """rank 0 adopts table `name` = {cx, cy} holding every rank's pairwise-complete rows; the others adopt it empty"""
    dev = torch.device("cuda", torch.cuda.current_device())
    cols = []
    for c in (cx, cy):
        b = ctx.column_buffers(table, c)
        if b["dtype"] not in (F.TG_INT64, F.TG_FLOAT64):
            raise NotImplementedError(f"multi-GPU Spearman over column '{table}.{c}': only Int64 / Float64 columns")
        n = b["n_rows"]
        v = _tensor_from_ptr(b["values"], n, dev) if n else torch.empty(0, dtype=torch.int64, device=dev)
        if b["dtype"] == F.TG_FLOAT64:
            v = v.view(torch.float64)
        cols.append((v, _unpack_validity(b["validity"], n, dev) if (b["validity"] and n) else None))
    mine = gather_pairs(complete_pairs(cols[0][0], cols[0][1], cols[1][0], cols[1][1]))
    m = mine.shape[0]
    xs = torch.zeros(m + 64, dtype=torch.float64, device=dev)
    ys = torch.zeros(m + 64, dtype=torch.float64, device=dev)
    xs[:m], ys[:m] = mine[:, 0], mine[:, 1]
    ctx.register_device_table(name, {cx: dict(dtype=F.TG_FLOAT64, n_rows=m, values=xs.data_ptr(), validity=None),
                                     cy: dict(dtype=F.TG_FLOAT64, n_rows=m, values=ys.data_ptr(), validity=None)},
                              keepalive=[xs, ys])


# ---------------------------------------------------------------------------------------------------------------------
# Distributed RANK() for Spearman: a sample sort around the engine's rank stages (tg_rank_*)
# ---------------------------------------------------------------------------------------------------------------------
RANK_SAMPLES_PER_SHARD = 2048


class GpuRankStages:
    """The device stages of the distributed rank computation, over the C ABI (include/termgpu.h tg_rank_*)."""

    def __init__(self, ctx):
        self.ctx, self.n, self.dev = ctx, 0, torch.device("cuda", torch.cuda.current_device())

    def begin(self, table, cx, cy):
        import ctypes as C
        n = C.c_int64()
        F.check(F.lib().tg_rank_begin(self.ctx.handle, table.encode(), cx.encode(), cy.encode(), C.byref(n)))
        self.n = n.value
        return self.n

    def local_sort(self):
        F.check(F.lib().tg_rank_local_sort(self.ctx.handle))

    def sample(self, m):
        import ctypes as C
        import numpy as np
        buf = (C.c_uint64 * max(m, 1))()
        got = F.check_slot(F.lib().tg_rank_sample(self.ctx.handle, m, buf))
        return np.frombuffer(buf, dtype=np.uint64, count=got).copy()

    def split(self, splitters, world):
        import ctypes as C
        import numpy as np
        sp = np.ascontiguousarray(splitters, dtype=np.uint64)
        counts = (C.c_int64 * world)()
        F.check(F.lib().tg_rank_split(self.ctx.handle, sp.ctypes.data_as(C.POINTER(C.c_uint64)), world, counts))
        return list(counts)

    def _views(self, kp, pp, n, payload_bytes):
        if n == 0:
            return (torch.empty(0, dtype=torch.int64, device=self.dev),
                    torch.empty(0, dtype=torch.int64 if payload_bytes == 8 else torch.int32, device=self.dev))
        return (_tensor_from_ptr(kp, n, self.dev, "<i8"), _tensor_from_ptr(pp, n, self.dev, "<i8" if payload_bytes == 8 else "<i4"))

    def send(self):
        import ctypes as C
        k, p, b = C.c_void_p(), C.c_void_p(), C.c_int32()
        F.check(F.lib().tg_rank_send_buffers(self.ctx.handle, C.byref(k), C.byref(p), C.byref(b)))
        self.payload_bytes = b.value
        return self._views(k.value, p.value, self.n, b.value)

    def recv(self, n_recv):
        import ctypes as C
        k, p = C.c_void_p(), C.c_void_p()
        F.check(F.lib().tg_rank_recv_buffers(self.ctx.handle, n_recv, C.byref(k), C.byref(p)))
        return self._views(k.value, p.value, n_recv, self.payload_bytes)

    def commit(self, n_recv):
        torch.cuda.current_stream().synchronize()  # the all-to-all wrote the receive buffers on torch's stream
        F.check(F.lib().tg_rank_recv_commit(self.ctx.handle, n_recv))
        self.n = n_recv

    def finish_x(self, base):
        F.check(F.lib().tg_rank_finish_x(self.ctx.handle, base))

    def finish_y(self, base, center):
        import ctypes as C
        n, sums = C.c_uint64(), (C.c_double * 5)()
        F.check(F.lib().tg_rank_finish_y(self.ctx.handle, base, center, C.byref(n), sums))
        return n.value, list(sums)

    def exchange_in_library(self):
        """tg_rank_exchange; None when the library cannot do it (no peer mapping): the caller runs the host-layer exchange"""
        import ctypes as C
        import time
        n, base, total, done = C.c_int64(), C.c_uint64(), C.c_uint64(), C.c_int32()
        sent0 = F.lib().tg_comm_bytes_sent(self.ctx.handle)
        t0 = time.perf_counter()
        F.check(F.lib().tg_rank_exchange(self.ctx.handle, C.byref(n), C.byref(base), C.byref(total), C.byref(done)))
        if not done.value:
            return None
        if PROFILE is not None:
            PROFILE["shuffle_ms"] += (time.perf_counter() - t0) * 1e3
            PROFILE["shuffle_bytes"] += F.lib().tg_comm_bytes_sent(self.ctx.handle) - sent0
        self.n = n.value
        return n.value, base.value, total.value

    def abort(self):
        F.lib().tg_rank_abort(self.ctx.handle)


def choose_splitters(samples, world):
    """world - 1 splitters at equal steps of the sorted union of every shard's samples (numpy uint64). Part p takes the
    keys in (splitter[p-1], splitter[p]]: a deterministic function of the gathered samples, identical on every rank."""
    import numpy as np
    s = np.sort(np.asarray(samples, dtype=np.uint64))
    if len(s) == 0:
        return np.zeros(world - 1, dtype=np.uint64)
    idx = [min(len(s) - 1, max(0, (p + 1) * len(s) // world - 1)) for p in range(world - 1)]
    return s[idx]


def _allgather_samples(samples, cap, dev):
    """every rank's sample keys (variable count <= cap) as one numpy uint64 array"""
    import numpy as np
    world = dist.get_world_size()
    mine = torch.zeros(cap + 1, dtype=torch.int64)
    mine[0] = len(samples)
    if len(samples):
        mine[1: 1 + len(samples)] = torch.from_numpy(samples.view(np.int64))
    mine = mine.to(dev)
    allv = [torch.zeros(cap + 1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allv, mine)
    out = []
    for t in allv:
        t = t.cpu().numpy()
        out.append(t[1: 1 + int(t[0])].view(np.uint64))
    return np.concatenate(out) if out else np.zeros(0, dtype=np.uint64)


def rank_sort_exchange(stages, dev):
    """One sample-sort exchange of the session's current (keys, payload): local sort -> splitters -> all-to-all by key
    range. Returns (n_recv, rank_base) with rank_base = the number of keys that went to lower ranks."""
    world, rank = dist.get_world_size(), dist.get_rank()
    prof = PROFILE if (PROFILE is not None and dev.type == "cuda") else None
    if isinstance(stages, GpuRankStages) and _ensure_comm(stages.ctx):
        got = stages.exchange_in_library()  # samples, splitters and the push of every pair to its range's owner: one call
        if got is not None:
            return got
    stages.local_sort()
    splitters = choose_splitters(_allgather_samples(stages.sample(RANK_SAMPLES_PER_SHARD), RANK_SAMPLES_PER_SHARD, dev), world)
    counts = stages.split(splitters, world)
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    send = torch.tensor(counts, dtype=torch.int64, device=dev)
    recv = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send)
    recv_counts = [int(x) for x in recv.tolist()]
    n_recv = sum(recv_counts)
    k_out, p_out = stages.send()
    k_in, p_in = stages.recv(n_recv)
    dist.all_to_all_single(k_in, k_out, output_split_sizes=recv_counts, input_split_sizes=counts)
    dist.all_to_all_single(p_in, p_out, output_split_sizes=recv_counts, input_split_sizes=counts)
    totals = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(totals, torch.tensor([n_recv], dtype=torch.int64, device=dev))
    totals = [int(t.item()) for t in totals]
    if prof is not None:
        ev1.record()
        ev1.synchronize()
        prof["shuffle_ms"] += ev0.elapsed_time(ev1)
        prof["shuffle_bytes"] += (8 + p_out.element_size()) * (sum(counts) - counts[rank])
    stages.commit(n_recv)
    return n_recv, sum(totals[:rank]), sum(totals)


def distributed_spearman(stages, table, cx, cy, dev):
    """Global minimum ranks of both columns over every rank's pairwise-complete rows, as this rank's partial state of the
    SPEARMAN aggregate: (u[8], f[8]) with u[0] = pairs that ended on this rank, f[0] = f[1] = (N + 1) / 2, f[2..6] = the
    sums of the shifted ranks, their squares and their product."""
    try:
        stages.begin(table, cx, cy)
        _, base, total = rank_sort_exchange(stages, dev)
        center = (total + 1.0) / 2.0
        stages.finish_x(base)
        _, base, _ = rank_sort_exchange(stages, dev)
        n, sums = stages.finish_y(base, center)
    except Exception:
        stages.abort()
        raise
    u, f = [0] * 8, [0.0] * 8
    if total >= 2:
        u[0], f[0], f[1] = n, center, center
        f[2:7] = sums
        if n == 0:  # an empty range contributes nothing (and must not carry a pivot of its own)
            f[0] = f[1] = 0.0
    elif rank_holds_all(n, total):
        u[0], f[0], f[1] = n, center, center
    return u, f


def rank_holds_all(n, total):
    return n == total and total > 0


def _adopt_fp_shard(ctx, name, recs: torch.Tensor):
    n = recs.numel() // 3
    vals = torch.zeros(n * 3 + 64, dtype=torch.int64, device=recs.device)
    vals[: n * 3] = recs
    ctx.register_device_table(name, {"tg_fp": dict(dtype=F.TG_FP128, n_rows=n, values=vals.data_ptr(), validity=None)}, keepalive=[vals])


def _register_empty_pair_table(ctx, name, cx, cy, dev):
    z = torch.zeros(64, dtype=torch.float64, device=dev)
    ctx.register_device_table(name, {cx: dict(dtype=F.TG_FLOAT64, n_rows=0, values=z.data_ptr(), validity=None),
                                     cy: dict(dtype=F.TG_FLOAT64, n_rows=0, values=z.data_ptr(), validity=None)}, keepalive=[z])


def _tensor_from_ptr(ptr, n, dev, typestr="<i8"):
    class _Wrap:  # __cuda_array_interface__ v3
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(w, device=dev)


def histogram_second_phase(plan, ctx, table):
    """Two-phase histogram (analyzers/advanced/histogram.rs:184-290: bucket bounds come from the table-wide MIN / MAX).
    After the merge every rank holds the global min / max; aggregates whose shards disagreed on the range are
    re-counted per shard against it, the counts are summed with one small all-reduce, installed, and the plan is
    finalized again. The pending list is derived from the merged state, so every rank takes the same branch."""
    pending = plan.histogram_pending()
    if not pending:
        return
    dev = _device()
    for i in pending:
        counts = plan.histogram_rebucket(ctx, table, i)
        # int64 on the wire (NCCL has no u64 sum in torch); counts are < 2^63
        t = torch.tensor(counts, dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        plan.histogram_install(i, [int(x) for x in t.tolist()])
    plan.finalize()


def execute_distributed(plan, ctx, table="data"):
    """Each rank: shuffle the keys of DISTINCT / FK aggregates, gather the column pairs of SPEARMAN aggregates,
    partial execute on its shard -> exchange -> ordered merge -> finalize."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        plan.execute(ctx, table)
        return
    aggs = plan.aggregates()
    if all(k in _SCAN_KINDS for k, _ in aggs):
        if PROFILE is not None:
            import time
            t0 = time.perf_counter()
        if execute_exchange_fused(plan, ctx, table):
            if PROFILE is not None:
                PROFILE["exchange_ms"] += 0.0  # inside the one call: not separable from the scan
            return
    temps, redirected, external = [], [], []
    try:
        for i, (kind, key) in enumerate(aggs):
            parts = key.split("|")
            if kind == KIND_DISTINCT:
                name = f"tg_shuffle_{i}_k"
                if len(parts) == 2 and _exact_shuffle_key(ctx, _column_dtype(ctx, table, parts[1])):
                    _shuffle_column(ctx, table, parts[1], name, by_range_if_dense=True)  # exact 64-bit keys
                else:
                    _shuffle_fingerprints(ctx, table, parts[1:], name)  # Utf8 / composite: 128-bit fingerprints
                temps.append(name)
                plan.redirect(i, 0, name)
                redirected.append((i, 0))
            elif kind == KIND_FK:
                (ct, cc), (pt, pc) = parts[1].split("."), parts[2].split(".")
                cname, pname = f"tg_shuffle_{i}_c", f"tg_shuffle_{i}_p"
                if _exact_shuffle_key(ctx, _column_dtype(ctx, ct, cc)) and _exact_shuffle_key(ctx, _column_dtype(ctx, pt, pc)):
                    _shuffle_column(ctx, ct, cc, cname)
                    temps.append(cname)
                    _shuffle_column(ctx, pt, pc, pname)
                    temps.append(pname)
                else:
                    # Utf8 keys: both sides travel as 128-bit fingerprint records (the identity the single-GPU path uses);
                    # counts are exact, the violation EXAMPLES (strings) are not reported across GPUs
                    _shuffle_fingerprints(ctx, ct, [cc], cname)
                    temps.append(cname)
                    _shuffle_fingerprints(ctx, pt, [pc], pname)
                    temps.append(pname)
                plan.redirect(i, 0, cname)
                plan.redirect(i, 1, pname)
                redirected += [(i, 0), (i, 1)]
            elif kind == KIND_SPEARMAN:
                name = f"tg_gather_{i}_s"
                if os.environ.get("TG_SPEARMAN_GATHER"):
                    _gather_spearman_columns(ctx, table, parts[1], parts[2], name)  # round-1 path: one GPU owns the pairs
                else:
                    # sample sort across the ranks; execute_partial then sees an empty table for this aggregate and the
                    # state computed here is installed afterwards
                    dev = torch.device("cuda", torch.cuda.current_device())
                    external.append((i, distributed_spearman(GpuRankStages(ctx), table, parts[1], parts[2], dev)))
                    _register_empty_pair_table(ctx, name, parts[1], parts[2], dev)
                temps.append(name)
                plan.redirect(i, 0, name)
                redirected.append((i, 0))
        plan.execute_partial(ctx, table)
        for i, (u, f) in external:
            plan.set_aggregate_partial(i, u, f)
        if PROFILE is not None:
            import time
            t0 = time.perf_counter()
        exchange_and_finalize(plan, ctx)
        histogram_second_phase(plan, ctx, table)
        if PROFILE is not None:
            PROFILE["exchange_ms"] += (time.perf_counter() - t0) * 1e3
    finally:
        for i, which in redirected:
            plan.redirect(i, which, None)
        for name in temps:
            ctx.deregister_table(name)
