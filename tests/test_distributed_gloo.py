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


# ---- string / length / KLL / grouped / histogram partials across ranks (SURVEY §8e K2, K4, K5 + the two-phase histogram) ----
M_ROWS = 6_007


def _mixed_table():
    rng = np.random.default_rng(21)
    s = []
    for i in range(M_ROWS):
        r = rng.random()
        s.append(None if r < 0.05 else (f"user{i}@example.com" if r < 0.6 else "x" * int(rng.integers(0, 30))))
    v = np.round(rng.normal(10.0, 4.0, M_ROWS), 2)
    v[:100] -= 50.0     # the global minimum lives in the first shard only
    v[-100:] += 80.0    # .. and the maximum in the last
    g = [f"g{int(x)}" for x in rng.integers(0, 7, M_ROWS)]
    g[M_ROWS - 1] = "only_in_last_shard"
    return pa.table({"s": pa.array(s, type=pa.string()), "v": pa.array(v, mask=rng.random(M_ROWS) < 0.1), "g": pa.array(g)})


def _hist_counts(vals, mn, mx, nb):
    """bucket rule of analyzers/advanced/histogram.rs:256-345 (first i with lower_i <= v < upper_i, else the last)"""
    rng_ = mx - mn
    w = rng_ / nb if (rng_ > 0.0 and nb > 1) else 1.0
    lowers = [mn + (i * w) for i in range(nb)]
    uppers = [(mx + w * 0.001) if i == nb - 1 else mn + ((i + 1) * w) for i in range(nb)]
    counts = [0] * nb
    for x in vals:
        b = nb - 1
        for i in range(nb):
            if lowers[i] <= x < uppers[i]:
                b = i
                break
        counts[b] += 1
    return counts


def mixed_shard_blob(plan, t):
    """partial blob of a row shard for REGEX (5), KLL (8), GROUPED (9), LENGTH (11), HIST (12) (+ the NUM aggregate the
    histogram reads its range from); layouts: term_b200/csrc/plan.hpp, kll_host.cpp, plan.cpp (grouped blob)"""
    from oracle import term_oracle as O
    cols = O.table_cols(t)
    n = t.num_rows
    out = [struct.pack("<Q", len(plan.aggregates()))]
    for kind, key in plan.aggregates():
        u, f, blob = [0] * 8, [0.0] * 8, b""
        parts = key.split("|")
        if kind == 5:
            flags = int(parts[2])
            m = O.regex_matches(cols[parts[1]], "|".join(parts[3:]), case_insensitive=bool(flags & 1), trim=bool(flags & 2))
            u[0], u[1], u[2] = sum(1 for x in m if x), sum(1 for x in m if x is None), n
        elif kind == 11:
            lo, hi = int(parts[2]), int(parts[3])
            c = cols[parts[1]]
            lens = [len(x) for x, ok in zip(c.values, c.valid) if ok]
            u[0], u[1], u[2] = sum(1 for L in lens if lo <= L <= hi), n - len(lens), n
        elif kind == 2:
            c = cols[parts[1]]
            vv = np.asarray(c.values, dtype=np.float64)[c.valid]
            u[0] = len(vv)
            if len(vv):
                K = float(vv[0])
                d = vv - K
                f[0], f[1], f[2], f[3], f[4], f[5] = K, math.fsum(d), math.fsum(d * d), float(vv.min()), float(vv.max()), math.fsum(vv)
        elif kind == 12:
            c = cols[parts[1]]
            vv = np.asarray(c.values, dtype=np.float64)[c.valid]
            nb = int(parts[2])
            if len(vv):
                f[0], f[1] = float(vv.min()), float(vv.max())
                blob = struct.pack(f"<{nb}Q", *_hist_counts(vv, f[0], f[1], nb))
        elif kind == 8:
            c = cols[parts[1]]
            vv = np.sort(np.asarray(c.values, dtype=np.float64)[c.valid])
            k = int(parts[2])
            u[0] = len(vv)
            mn, mx = (float(vv[0]), float(vv[-1])) if len(vv) else (math.inf, -math.inf)
            blob = struct.pack("<QddQQQ", len(vv), mn, mx, k, max(8 * k, 64), 1) + struct.pack("<Q", len(vv)) + vv.tobytes()
        elif kind == 9:
            tgt, gcols = cols[parts[1]], [cols[p] for p in parts[2:]]
            groups = {}
            for i in range(n):
                gk = "\x1f".join(str(gc.values[i]) for gc in gcols)
                tt, nn = groups.get(gk, (0, 0))
                groups[gk] = (tt + 1, nn + (1 if tgt.valid[i] else 0))
            blob = struct.pack("<Q", len(groups))
            for gk, (tt, nn) in groups.items():
                kb = gk.encode()
                blob += struct.pack("<I", len(kb)) + kb + struct.pack("<QQ", tt, nn)
        else:
            raise AssertionError(f"unexpected aggregate {kind} {key}")
        pad = (-len(blob)) % 8
        out.append(struct.pack("<QQ", kind, 0) + struct.pack("<8Q", *u) + struct.pack("<8d", *f) + struct.pack("<Q", len(blob)) + blob +
                   b"\0" * pad + struct.pack("<Q", 0))
    return b"".join(out)


def _mixed_plan(T):
    cb = (T.Check.builder("c").validates_regex("s", "@", 0.5).validates_email("s", 0.5).has_min_length("s", 5).has_length_between("s", 1, 20))
    suite = T.ValidationSuite.builder("mixed").check(cb.build()).build()
    plan, slots = suite.build_plan()
    kll = T.KllSketchAnalyzer("v", k=64, quantiles=[0.1, 0.5, 0.9])._add_to(plan)
    grp = T.GroupedCompletenessAnalyzer("v", ["g"])._add_to(plan)
    hist = T.HistogramAnalyzer("v", 12)._add_to(plan)
    return plan, slots, kll, grp, hist


def mixed_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import term_b200 as T
        from term_b200.distributed import allgather_blobs, merge_partials
        t = _mixed_table()
        bounds = [0, M_ROWS // 3, M_ROWS // 3, M_ROWS][: world + 1] if world == 3 else [M_ROWS * r // world for r in range(world + 1)]
        sh = t.slice(bounds[rank], bounds[rank + 1] - bounds[rank])  # world 3: the middle rank holds an EMPTY shard
        plan, slots, kll, grp, hist = _mixed_plan(T)
        assert sorted({k for k, _ in plan.aggregates()}) == [2, 5, 8, 9, 11, 12]
        merge_partials(plan, allgather_blobs(mixed_shard_blob(plan, sh)))
        # two-phase histogram: the merged NUM aggregate holds the global range; every rank re-counts its shard, the
        # counts are all-reduced (term_b200.distributed.histogram_second_phase does this with the device re-count)
        pending = plan.histogram_pending()
        assert len(pending) == 1
        assert plan.analyzer_result(hist).error == 2
        vv = np.asarray(sh.column("v").drop_null())
        full = np.asarray(t.column("v").drop_null())
        mine = torch.tensor(_hist_counts(vv, float(full.min()), float(full.max()), 12), dtype=torch.int64)
        dist.all_reduce(mine)
        plan.histogram_install(pending[0], [int(x) for x in mine.tolist()])
        plan.finalize()
        res = [plan.result(s) for _, _, s in slots]
        k, g, h = plan.analyzer_result(kll), plan.analyzer_result(grp), plan.analyzer_result(hist)
        q.put((rank, ([(r.name, r.status.name, r.metric, r.message) for r in res], k.map, k.u[0], g.map, h.map, plan.kll_levels(kll))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_string_kll_grouped_histogram_partials_merge(built_lib, world):
    from oracle import term_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=mixed_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(1, world):
        assert results[r] == results[0], "ranks disagree after the ordered merge"
    res, kmap, kn, gmap, hmap, levels = results[0]
    t = _mixed_table()
    want = [O.format_constraint(t, "s", "Regex", 0.5, arg="@"), O.format_constraint(t, "s", "Email", 0.5),
            O.length_constraint(t, "s", "Min", 5), O.length_constraint(t, "s", "Between", 1, 20)]
    for (name, status, metric, message), o in zip(res, want):
        assert status.lower() == o.status and metric == o.metric and message == o.message, (name, metric, o)
    # KLL: 5 400 values through a k = 64 sketch (capacity 512): compacted on merge, still inside the reference bound
    clean = np.sort(np.asarray(t.column("v").drop_null()))
    assert kn == len(clean) and kmap["count"] == len(clean) and kmap["min"] == clean[0] and kmap["max"] == clean[-1]
    for qq in (0.1, 0.5, 0.9):
        assert O.rank_error(clean, kmap["quantile_" + O.rust_f64(qq)], qq) <= 1.65 / math.sqrt(64)
    assert sum(len(l) for l in levels) <= 512 and len(levels) >= 4, "the merged sketch is a compacted multi-level ladder"
    total_w = sum(len(l) << i for i, l in enumerate(levels))
    assert total_w == len(clean), "weight-preserving compaction: item weights add up to the exact count"
    for l in levels:
        assert l == sorted(l)
    # grouped: a group only one rank saw survives the merge
    gw = O.grouped_completeness(t, "v", ["g"])
    assert {k: v for k, v in gmap.items() if not k.startswith("__")} == {k[0]: nn / tt for k, (tt, nn) in gw.items()}
    assert "only_in_last_shard" in gmap
    hw = O.an_histogram(t, "v", 12)
    assert hmap["min"] == hw["min"] and hmap["max"] == hw["max"] and hmap["total_count"] == hw["total_count"]
    for i, (lo, hi, cnt) in enumerate(hw["buckets"]):
        assert hmap[f"bucket_{i}.lower"] == lo and hmap[f"bucket_{i}.upper"] == hi and hmap[f"bucket_{i}.count"] == cnt


# ---- Spearman as a distributed sample sort: the host orchestration over gloo, the device stages in numpy ----
class NumpyRankStages:
    """what term_b200.distributed.GpuRankStages does through tg_rank_* on the device, restated in numpy"""

    def __init__(self, shard):
        self.shard = shard

    @staticmethod
    def _key(v):
        v = np.asarray(v, dtype=np.float64) + 0.0  # -0.0 -> +0.0
        b = v.view(np.uint64)
        return np.where(b >> np.uint64(63), ~b, b | np.uint64(1 << 63))

    def begin(self, table, cx, cy):
        x, y = self.shard.column(cx), self.shard.column(cy)
        ok = np.asarray(x.is_valid()) & np.asarray(y.is_valid())
        self.k = self._key(np.asarray(x.fill_null(0)).astype(np.float64)[ok])
        self.p = self._key(np.asarray(y.fill_null(0)).astype(np.float64)[ok])
        self.phase = 0
        return len(self.k)

    def local_sort(self):
        o = np.argsort(self.k, kind="stable")
        self.k, self.p = self.k[o], self.p[o]

    def sample(self, m):
        n = len(self.k)
        m = min(m, n)
        return self.k[[((2 * i + 1) * n) // (2 * m) for i in range(m)]] if m else np.zeros(0, dtype=np.uint64)

    def split(self, splitters, world):
        pos = [int(np.searchsorted(self.k, s, side="right")) for s in splitters] + [len(self.k)]
        out, prev = [], 0
        for q in pos:
            q = max(q, prev)
            out.append(q - prev)
            prev = q
        return out

    def send(self):
        import torch
        pd = np.int64 if self.phase == 0 else np.int32
        return torch.from_numpy(self.k.view(np.int64).copy()), torch.from_numpy(self.p.view(pd).copy() if self.phase == 0 else self.p.astype(np.uint32).view(np.int32).copy())

    def recv(self, n_recv):
        import torch
        self.rk = torch.zeros(n_recv, dtype=torch.int64)
        self.rp = torch.zeros(n_recv, dtype=torch.int64 if self.phase == 0 else torch.int32)
        return self.rk, self.rp

    def commit(self, n_recv):
        self.k = self.rk.numpy().view(np.uint64).copy()
        self.p = self.rp.numpy().view(np.uint64 if self.phase == 0 else np.uint32).copy()

    @staticmethod
    def _min_ranks(sorted_keys, base):
        n = len(sorted_keys)
        head = np.ones(n, dtype=bool)
        head[1:] = sorted_keys[1:] != sorted_keys[:-1]
        pos = np.where(head, np.arange(n), 0)
        return base + np.maximum.accumulate(pos) + 1

    def finish_x(self, base):
        self.local_sort()
        rx = self._min_ranks(self.k, base)
        self.k, self.p, self.phase = self.p, rx.astype(np.uint32), 1

    def finish_y(self, base, center):
        self.local_sort()
        ry = self._min_ranks(self.k, base).astype(np.float64) - center
        rx = self.p.astype(np.float64) - center
        return len(self.k), [math.fsum(rx), math.fsum(ry), math.fsum(rx * rx), math.fsum(ry * ry), math.fsum(rx * ry)]

    def abort(self):
        pass


def _tie_table():
    rng = np.random.default_rng(13)
    n = 9_001
    x = np.round(rng.normal(0, 3, n), 0)     # heavy ties
    y = np.round(0.5 * x + rng.normal(0, 2, n), 1)
    x[:50] = 7.0                              # one value that dominates a whole sample range
    return pa.table({"x": pa.array(x, mask=rng.random(n) < 0.1), "y": pa.array(y, mask=rng.random(n) < 0.05)})


def sample_sort_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import term_b200 as T
        from term_b200.distributed import allgather_blobs, distributed_spearman, merge_partials
        t = _tie_table()
        n = t.num_rows
        bounds = [0, n // 4, n // 4, n][: world + 1] if world == 3 else [n * r // world for r in range(world + 1)]
        shard = t.slice(bounds[rank], bounds[rank + 1] - bounds[rank])  # world 3: the middle rank starts EMPTY
        u, f = distributed_spearman(NumpyRankStages(shard), "data", "x", "y", torch.device("cpu"))
        plan = T.Plan()
        slot = T.CorrelationAnalyzer.spearman("x", "y")._add_to(plan)
        blob = struct.pack("<Q", 1) + struct.pack("<QQ", 10, 0) + struct.pack("<8Q", *u) + struct.pack("<8d", *f) + struct.pack("<QQ", 0, 0)
        merge_partials(plan, allgather_blobs(blob))
        r = plan.analyzer_result(slot)
        q.put((rank, (r.u[0], r.metric_double, u[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_spearman_sample_sort_gives_global_ranks(built_lib, world):
    from oracle import term_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 37500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=sample_sort_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = _tie_table()
    x, _ = O.pair_values(t, "x", "y")
    want = O.an_correlation(t, "x", "y", "spearman")
    assert sum(results[r][2] for r in range(world)) == len(x), "every pair ends on exactly one rank"
    for r in range(world):
        assert results[r][0] == len(x)
        assert abs(results[r][1] - want) <= 1e-9, (results[r][1], want)
