"""N>1 host path on CPU: world_size-2 `gloo` ranks each hold a row shard, assemble the shard's partial
aggregate blob (here from numpy — on a GPU box tg_plan_execute_partial produces it), all-gather the
blobs with term_b200.distributed, merge them in rank order through the C ABI and finalize. Every rank must
reproduce the oracle's single-table answer (AnalyzerState::merge semantics, analyzers/traits.rs:154-179)."""
import math
import os
import struct

import numpy as np
import pyarrow as pa
import pytest
import torch.multiprocessing as mp

N_ROWS = 20_001


def make_table():
    rng = np.random.default_rng(99)
    x = rng.normal(50.0, 7.0, N_ROWS)
    y = 0.3 * x + rng.normal(0, 2.0, N_ROWS)
    k = rng.integers(-2**62, 2**62, N_ROWS)
    return pa.table({"x": pa.array(x, mask=rng.random(N_ROWS) < 0.1), "y": pa.array(y, mask=rng.random(N_ROWS) < 0.1),
                     "k": pa.array(k, mask=rng.random(N_ROWS) < 0.1)})


def build_plan(T):
    A = T.Assertion
    cb = (T.Check.builder("c").has_size(A.Equals(float(N_ROWS))).completeness("x", 0.95)
          .has_mean("x", A.Between(49.0, 51.0)).has_standard_deviation("x", A.LessThan(100.0))
          .has_min("x", A.GreaterThan(-1e9)).has_max("x", A.LessThan(1e9)).has_sum("k", A.LessThan(1e300))
          .has_min("k", A.LessThan(0.0)).has_correlation("x", "y", A.GreaterThan(0.5)).satisfies("x > 45", "x must exceed 45"))
    suite = T.ValidationSuite.builder("dist").check(cb.build()).build()
    return suite.build_plan()


def shard_blob(plan, t):
    """Partial blob of one row shard, layout documented in term_b200/csrc/plan.cpp (partial_export)."""
    cols = {n: (np.asarray(t.column(n).fill_null(0)), np.asarray(t.column(n).is_valid())) for n in t.schema.names}
    n = t.num_rows
    out = [struct.pack("<Q", len(plan.aggregates()))]
    for kind, key in plan.aggregates():
        u, f = [0] * 8, [0.0] * 8
        parts = key.split("|")
        if kind == 0:
            u[0] = n
        elif kind == 1:
            u[0], u[1] = n, int(cols[parts[1]][1].sum())
        elif kind == 2:
            v, ok = cols[parts[1]]
            vv = v[ok]
            u[0] = len(vv)
            if len(vv):
                K = float(vv[0])
                d = vv.astype(np.float64) - K
                f[0], f[1], f[2] = K, math.fsum(d), math.fsum(d * d)
                f[5] = math.fsum(vv.astype(np.float64))
                if v.dtype == np.int64:
                    u[1] = int(sum(int(a) for a in vv)) & (2**64 - 1)
                    u[2], u[3], u[4] = int(vv.min()) & (2**64 - 1), int(vv.max()) & (2**64 - 1), 1
                else:
                    f[3], f[4] = float(vv.min()), float(vv.max())
        elif kind == 3:
            (vx, okx), (vy, oky) = cols[parts[1]], cols[parts[2]]
            m = okx & oky
            a, b = vx[m].astype(np.float64), vy[m].astype(np.float64)
            u[0] = int(m.sum())
            if u[0]:
                Kx, Ky = float(a[0]), float(b[0])
                dx, dy = a - Kx, b - Ky
                f[0], f[1] = Kx, Ky
                f[2], f[3], f[4], f[5], f[6] = math.fsum(dx), math.fsum(dy), math.fsum(dx * dx), math.fsum(dy * dy), math.fsum(dx * dy)
        elif kind == 4:
            v, ok = cols["x"]
            u[0], u[2] = int(((v > 45) & ok).sum()), n
        else:
            raise AssertionError(f"unexpected aggregate {kind} {key}")
        out.append(struct.pack("<QQ", kind, 0) + struct.pack("<8Q", *u) + struct.pack("<8d", *f) + struct.pack("<QQ", 0, 0))
    return b"".join(out)


def worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import term_b200 as T
        from term_b200.distributed import allgather_blobs, allgather_blobs_fixed, merge_partials
        t = make_table()
        lo, hi = N_ROWS * rank // world, N_ROWS * (rank + 1) // world
        plan, slots = build_plan(T)
        mine = shard_blob(plan, t.slice(lo, hi - lo))
        blobs = allgather_blobs(mine)
        assert len(blobs) == world
        # the one-collective path used for plans of fixed-size aggregates must deliver the same bytes; a blob that
        # does not fit makes every rank fall back (None)
        assert allgather_blobs_fixed(mine, 8192) == blobs
        assert allgather_blobs_fixed(mine, 64) is None
        merge_partials(plan, blobs)
        res = [plan.result(s) for _, _, s in slots]
        q.put((rank, [(r.name, r.status.name, r.metric, r.message) for r in res]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_merge_matches_oracle(built_lib):
    from oracle import term_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0] == results[1], "ranks disagree after the ordered merge"
    t = make_table()
    want = [O.size(t, ("Equals", float(N_ROWS))), O.completeness(t, "x", 0.95), O.statistic(t, "x", "Mean", ("Between", 49.0, 51.0)),
            O.statistic(t, "x", "StandardDeviation", ("LessThan", 100.0)), O.statistic(t, "x", "Min", ("GreaterThan", -1e9)),
            O.statistic(t, "x", "Max", ("LessThan", 1e9)), O.statistic(t, "k", "Sum", ("LessThan", 1e300)),
            O.statistic(t, "k", "Min", ("LessThan", 0.0)), O.correlation(t, "x", "y", "Pearson", ("GreaterThan", 0.5)),
            O.custom_sql(t, "x > 45", "x must exceed 45")]
    exact = {"size", "completeness", "min", "max", "sum", "custom_sql"}
    for (name, status, metric, message), o in zip(results[0], want):
        assert status.lower() == o.status, (name, status, o)
        if name in exact:
            assert metric == o.metric and message == o.message, (name, metric, o)
        else:
            assert abs(metric - o.metric) <= 1e-9 * max(1.0, abs(o.metric)), (name, metric, o.metric)


def test_partial_blob_roundtrip_and_rejects_foreign_blobs(built_lib):
    import term_b200 as T
    t = make_table()
    plan, _ = build_plan(T)
    plan.partial_reset()
    plan.partial_merge(shard_blob(plan, t))
    blob = plan.partial_export()
    other, _ = build_plan(T)
    other.partial_reset()
    other.partial_merge(blob)
    assert other.partial_export() == blob
    small = T.Plan()
    T.SizeConstraint(T.Assertion.Equals(1.0))._add_to(small)
    with pytest.raises(T.TermGpuError):
        small.partial_merge(blob)
    with pytest.raises(T.TermGpuError):
        plan.partial_merge(blob[: len(blob) // 2])


# ---- hash shuffle for uniqueness / foreign key: all-to-all over gloo, states add up exactly ----
def _shuffle_table():
    rng = np.random.default_rng(5)
    n = 30_011
    k = rng.integers(0, n // 2, n)
    k[rng.random(n) < 0.01] = -1
    return pa.table({"k": pa.array(k, mask=rng.random(n) < 0.05)})


def _distinct_blob(plan, keys, n_nulls):
    """partial blob of a hash shard: A_DISTINCT state u0 rows u1 distinct u2 singleton groups u3/u4 NULL rows
    u5 distinct with NULL as a value (term_b200/csrc/plan.hpp)"""
    aggs = plan.aggregates()
    assert [k for k, _ in aggs] == [6]
    _, counts = np.unique(keys, return_counts=True)
    d, singles = len(counts), int((counts == 1).sum())
    u = [len(keys) + n_nulls, d, singles + (1 if n_nulls == 1 else 0), n_nulls, n_nulls, d + (1 if n_nulls else 0), 0, 0]
    return (struct.pack("<Q", 1) + struct.pack("<QQ", 6, 0) + struct.pack("<8Q", *u) + struct.pack("<8d", *([0.0] * 8)) +
            struct.pack("<QQ", 0, 0))


def shuffle_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import term_b200 as T
        from term_b200.distributed import allgather_blobs, merge_partials, shuffle_keys
        from tests.helpers import hash_rank_np
        t = _shuffle_table()
        n = t.num_rows
        shard = t.slice(n * rank // world, n * (rank + 1) // world - n * rank // world)
        vals = np.asarray(shard.column("k").fill_null(0))
        ok = np.asarray(shard.column("k").is_valid())
        keys = vals[ok]
        dest = hash_rank_np(keys, world)  # what tg_table_partition_keys computes on the device
        order = np.argsort(dest, kind="stable")
        counts = [int((dest == r).sum()) for r in range(world)]
        mine, my_nulls = shuffle_keys(torch.from_numpy(keys[order].copy()), counts, int((~ok).sum()))
        mine = mine.numpy()
        assert (hash_rank_np(mine, world) == rank).all(), "received a key that hashes to another rank"
        plan = T.Plan()
        slot = T.UniquenessConstraint(["k"], T.UniquenessType.UniqueValueRatio, assertion=T.Assertion.GreaterThan(0.0))._add_to(plan)
        slot2 = T.DistinctnessAnalyzer("k")._add_to(plan)
        blobs = allgather_blobs(_distinct_blob(plan, mine, my_nulls))
        merge_partials(plan, blobs)
        r = plan.result(slot)
        a = plan.analyzer_result(slot2)
        q.put((rank, (r.status.name, r.metric, r.message, a.u[:2], a.metric_double, len(mine))))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_hash_shuffle_uniqueness(built_lib):
    from oracle import term_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=shuffle_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][:5] == results[1][:5], "ranks disagree after the merge"
    t = _shuffle_table()
    o = O.uniqueness(t, ["k"], "UniqueValueRatio", 1.0, ("GreaterThan", 0.0))
    nn, d, m = O.an_distinctness(t, "k")
    status, metric, message, u, md, _ = results[0]
    assert status.lower() == o.status and metric == o.metric and message == o.message
    assert u == [nn, d] and md == m
    assert results[0][5] + results[1][5] == nn, "every valid key lands on exactly one rank"
    assert min(results[0][5], results[1][5]) > nn // 4, "the hash split is badly skewed"


# ---- Spearman across ranks: global ranks need all pairwise-complete rows on one rank ----
def spearman_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from scipy.stats import rankdata
        import term_b200 as T
        from term_b200.distributed import allgather_blobs, complete_pairs, gather_pairs, merge_partials
        t = make_table()
        # uneven shards, the last one empty: the gather pads to the largest
        bounds = [0, N_ROWS // 3, N_ROWS, N_ROWS][: world + 1] if world == 3 else [N_ROWS * r // world for r in range(world + 1)]
        sh = t.slice(bounds[rank], bounds[rank + 1] - bounds[rank])
        col = lambda n: (torch.from_numpy(np.asarray(sh.column(n).fill_null(0)).copy()), torch.from_numpy(np.asarray(sh.column(n).is_valid()).copy()))  # noqa: E731
        (x, xv), (k, kv) = col("x"), col("k")
        pairs = gather_pairs(complete_pairs(x, xv, k, kv))  # f64 x, i64 k -> DOUBLE
        plan = T.Plan()
        slot = T.CorrelationAnalyzer.spearman("x", "k")._add_to(plan)
        (kind, key), = plan.aggregates()
        assert kind == 10 and key == "spearman|x|k"
        u, f = [0] * 8, [0.0] * 8
        if rank == 0:  # what exec_spearman_job computes on the gathered table: co-moments of the min-ranks
            a = rankdata(pairs[:, 0].numpy(), method="min").astype(np.float64)
            b = rankdata(pairs[:, 1].numpy(), method="min").astype(np.float64)
            u[0] = len(a)
            da, db = a - a[0], b - b[0]
            f[0], f[1] = float(a[0]), float(b[0])
            f[2], f[3], f[4], f[5], f[6] = math.fsum(da), math.fsum(db), math.fsum(da * da), math.fsum(db * db), math.fsum(da * db)
        else:
            assert pairs.shape == (0, 2)  # the other ranks hold nothing and contribute an empty partial
        blob = struct.pack("<Q", 1) + struct.pack("<QQ", kind, 0) + struct.pack("<8Q", *u) + struct.pack("<8d", *f) + struct.pack("<QQ", 0, 0)
        merge_partials(plan, allgather_blobs(blob))
        r = plan.analyzer_result(slot)
        q.put((rank, (r.u[0], r.metric_double, int(pairs.shape[0]))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_spearman_gather_gives_global_ranks(built_lib, world):
    from oracle import term_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=spearman_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = make_table()
    x, k = O.pair_values(t, "x", "k")
    want = O.an_correlation(t, "x", "k", "spearman")
    assert results[0][2] == len(x) and all(results[r][2] == 0 for r in range(1, world))
    for r in range(world):
        assert results[r][0] == len(x), "every rank finalizes to the global pair count"
        assert abs(results[r][1] - want) <= 1e-9, (results[r][1], want)
