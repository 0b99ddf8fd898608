"""Row-shard merge parity: for every aggregate family whose multi-GPU path is "evaluate the shard, exchange
the partial state, merge in rank order", a table is cut into row shards (one of them empty), each shard goes through
tg_plan_execute_partial -> tg_plan_partial_export, the blobs are merged IN ORDER into a fresh state
(tg_plan_partial_reset / _merge / _finalize) and the result must equal (a) the one-pass result on the whole table and
(b) the oracle. This is the contract of AnalyzerState::merge (analyzers/traits.rs:154-179),
GroupedCompletenessState::merge (analyzers/basic/grouped_completeness.rs:37-84) and KllSketch::merge
(analyzers/advanced/kll_sketch.rs:327-366). Counts, ratios and messages bit-exact; KLL: count / min / max exact,
quantiles inside the reference's rank-error bound (and exact below the sketch capacity)."""
import math

import numpy as np
import pyarrow as pa
import pytest

import term_b200 as T
from oracle import term_oracle as O

pytestmark = pytest.mark.gpu


def _cuts(n, fracs=(0.0, 0.37, 0.37, 0.8, 1.0)):
    """row boundaries of the shards: the second shard is EMPTY, the others ragged (not multiples of 64)"""
    return [int(n * f) for f in fracs]


def _register_shards(ctx, prefix, table, cuts):
    names = []
    for i in range(len(cuts) - 1):
        name = f"{prefix}_{i}"
        ctx.register_table(name, table.slice(cuts[i], cuts[i + 1] - cuts[i]))
        names.append(name)
    return names


def _merged(ctx, plan, shard_names, order=None):
    """partials of every shard -> ordered merge -> (two-phase histogram) -> finalize"""
    blobs = []
    for nm in shard_names:
        plan.execute_partial(ctx, nm)
        blobs.append(plan.partial_export())
    plan.partial_reset()
    for k in (order or range(len(blobs))):
        plan.partial_merge(blobs[k])
    plan.finalize()
    for i in plan.histogram_pending():  # what distributed.histogram_second_phase does with an all-reduce
        total = None
        for nm in shard_names:
            c = plan.histogram_rebucket(ctx, nm, i)
            total = c if total is None else [a + b for a, b in zip(total, c)]
        plan.histogram_install(i, total)
    plan.finalize()
    return blobs


def _strings(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        r = rng.random()
        if r < 0.03:
            out.append(None)
        elif r < 0.55:
            out.append(f"user{int(rng.integers(0, 10**6))}@example{int(rng.integers(0, 50))}.com")
        elif r < 0.65:
            out.append("%03d-%02d-%04d" % (rng.integers(1, 899), rng.integers(1, 99), rng.integers(1, 9999)))
        elif r < 0.75:
            out.append(" ".join("%04d" % rng.integers(0, 9999) for _ in range(4)))
        elif r < 0.8:
            out.append("héllo wörld 你好 🦀"[: int(rng.integers(0, 16))])
        else:
            out.append("".join(rng.choice(list("abc XYZ@.-_1290"), int(rng.integers(0, 40)))))
    return out


@pytest.mark.parametrize("n", [5, 1000, 70_000])
def test_string_and_length_partials_merge_to_the_one_pass_answer(ctx, n):
    t = pa.table({"s": pa.array(_strings(n, n), type=pa.string())})
    cuts = _cuts(n)
    names = _register_shards(ctx, f"mg_str_{n}", t, cuts)
    whole = f"mg_str_{n}_all"
    ctx.register_table(whole, t)
    try:
        cb = (T.Check.builder("pii").validates_regex("s", "@", 0.5).validates_email("s", 0.5).contains_ssn("s", 0.05)
              .validates_credit_card("s", 0.2, True)
              .has_format("s", T.FormatType.Regex, 0.1, T.FormatOptions.lenient(), arg=r"^[A-Z]+@")
              .has_min_length("s", 3).has_max_length("s", 30).has_length_between("s", 2, 25).has_exact_length("s", 11).is_not_empty("s")
              .constraint(T.DataTypeConstraint("s", T.DataType.Integer, 0.0)))
        suite = T.ValidationSuite.builder("s").table_name(whole).check(cb.build()).build()
        plan, slots = suite.build_plan()
        kinds = {k for k, _ in plan.aggregates()}
        assert 5 in kinds and 11 in kinds  # REGEX and LENGTH aggregates are what this test is about
        plan.execute(ctx, whole)
        one_pass = [plan.result(s) for _, _, s in slots]
        _merged(ctx, plan, names)
        merged = [plan.result(s) for _, _, s in slots]
        want = [O.format_constraint(t, "s", "Regex", 0.5, arg="@"), O.format_constraint(t, "s", "Email", 0.5),
                O.format_constraint(t, "s", "SocialSecurityNumber", 0.05, trim=True),
                O.format_constraint(t, "s", "CreditCard", 0.2, flag=True),
                O.format_constraint(t, "s", "Regex", 0.1, arg=r"^[A-Z]+@", case_sensitive=False, trim=True),
                O.length_constraint(t, "s", "Min", 3), O.length_constraint(t, "s", "Max", 30),
                O.length_constraint(t, "s", "Between", 2, 25), O.length_constraint(t, "s", "Exactly", 11),
                O.length_constraint(t, "s", "NotEmpty"), O.data_type(t, "s", "Integer", 0.0)]
        for m, g, o in zip(merged, one_pass, want):
            assert (m.status, m.metric, m.message) == (g.status, g.metric, g.message), (m, g)
            assert m.status.name.lower() == o.status and m.metric == o.metric and m.message == o.message, (m, o)
        # the merge is a sum: any order of the same blobs gives the same state
        _merged(ctx, plan, names, order=[3, 0, 2, 1])
        assert [(r.status, r.metric, r.message) for r in (plan.result(s) for _, _, s in slots)] == \
               [(r.status, r.metric, r.message) for r in merged]
    finally:
        for nm in names + [whole]:
            ctx.deregister_table(nm)


@pytest.mark.parametrize("n,k", [(9, 64), (3000, 256), (2_000_000, 256)])
def test_kll_partials_merge_within_the_reference_bound(ctx, n, k):
    """KllSketch::merge (kll_sketch.rs:327-366): the merged sketch keeps count / min / max exact and answers every
    quantile inside the 1.65/sqrt(k) rank-error bound (:397-399); below the capacity it is the exact order statistic"""
    rng = np.random.default_rng(n)
    vals = rng.lognormal(0.0, 1.5, n)
    vals[: n // 3] += 50.0  # shards with different distributions: the merge must weigh them by their counts
    mask = rng.random(n) < 0.07
    if n > 100:
        vals[5] = np.nan
    t = pa.table({"x": pa.array(vals, mask=mask)})
    cuts = _cuts(n)
    names = _register_shards(ctx, f"mg_kll_{n}", t, cuts)
    try:
        qs = [0.0, 0.01, 0.25, 0.5, 0.75, 0.95, 0.99, 1.0]
        plan = T.Plan()
        slot = T.KllSketchAnalyzer("x", k=k, quantiles=qs)._add_to(plan)
        assert 8 in {kk for kk, _ in plan.aggregates()}
        _merged(ctx, plan, names)
        r = plan.analyzer_result(slot)
        clean = np.sort(vals[~mask & ~np.isnan(vals)])
        assert r.error == 0 and r.u[0] == len(clean) and r.map["count"] == len(clean)
        assert r.map["min"] == clean[0] and r.map["max"] == clean[-1]
        bound = 1.65 / math.sqrt(k)
        prev = -math.inf
        for q in qs:
            est = r.map["quantile_" + O.rust_f64(q)]
            assert clean[0] <= est <= clean[-1] and est >= prev
            prev = est
            assert O.rank_error(clean, est, q) <= bound, (q, est)
            if n <= 2 * k:
                target = max(1, math.ceil(q * len(clean)))
                assert est == (clean[0] if q == 0.0 else clean[-1] if q == 1.0 else clean[target - 1])
    finally:
        for nm in names:
            ctx.deregister_table(nm)


def test_grouped_completeness_partials_merge_exactly(ctx):
    """GroupedCompletenessState::merge (grouped_completeness.rs:37-84): per-group (total, non-null) counts add; groups that
    only some shards saw survive; Utf8, Int64 (with NULL keys) and two-column groupings"""
    rng = np.random.default_rng(11)
    n = 120_000
    region = np.array([f"region{v}" for v in rng.integers(0, 16, n)], dtype=object)
    region[: n // 2][region[: n // 2] == "region3"] = "region0"  # region3 only exists in the later shards
    cat = [f"cat{v}" for v in rng.integers(0, 200, n)]
    gid = rng.integers(0, 3000, n)
    t = pa.table({"g1": pa.array(region.tolist()), "g2": pa.array(cat), "gid": pa.array(gid, mask=rng.random(n) < 0.01),
                  "v": pa.array(rng.normal(0, 1, n), mask=rng.random(n) < 0.2)})
    names = _register_shards(ctx, "mg_grp", t, _cuts(n))
    ctx.register_table("mg_grp_all", t)
    try:
        for groups in (["g1"], ["g2"], ["g1", "g2"], ["gid"]):
            plan = T.Plan()
            slot = T.GroupedCompletenessAnalyzer("v", groups)._add_to(plan)
            assert 9 in {k for k, _ in plan.aggregates()}
            plan.execute(ctx, "mg_grp_all")
            one = plan.analyzer_result(slot).map
            _merged(ctx, plan, names)
            got = plan.analyzer_result(slot).map
            assert got == one
            want = O.grouped_completeness(t, "v", groups)

            def text(x):
                return "NULL" if x is None else str(x)

            assert {k: v for k, v in got.items() if not k.startswith("__")} == \
                   {"_".join(text(x) for x in k): nn / tt for k, (tt, nn) in want.items()}
            assert got["__overall__"] == sum(nn for _, nn in want.values()) / sum(tt for tt, _ in want.values())
    finally:
        for nm in names + ["mg_grp_all"]:
            ctx.deregister_table(nm)


def _hist_maps_equal(a, b):
    assert a.keys() == b.keys()
    for k in a:
        if k in ("mean", "std_dev", "sum", "sum_squared"):  # float sums: shard order changes the rounding
            assert a[k] == pytest.approx(b[k], rel=1e-9, abs=1e-9), k
        else:
            assert a[k] == b[k], (k, a[k], b[k])


@pytest.mark.parametrize("n,nb", [(4, 3), (5000, 10), (150_000, 1000)])
def test_histogram_shards_merge_through_the_global_range(ctx, n, nb):
    """histogram.rs:184-290: bucket bounds come from the table-wide MIN / MAX, so shards with different local ranges
    need the second phase (pending -> rebucket against the merged min / max -> summed counts installed)"""
    rng = np.random.default_rng(n + nb)
    vals = np.round(rng.normal(50.0, 20.0, n), 1)
    vals[-1] = 400.0  # the global maximum sits in the last shard, the minimum in the first
    vals[0] = -300.0
    t = pa.table({"v": pa.array(vals, mask=(rng.random(n) < 0.1) & (np.arange(n) > 0) & (np.arange(n) < n - 1))})
    names = _register_shards(ctx, f"mg_hist_{n}", t, _cuts(n))
    whole = f"mg_hist_{n}_all"
    ctx.register_table(whole, t)
    try:
        plan = T.Plan()
        slot = T.HistogramAnalyzer("v", nb)._add_to(plan)
        hist_aggs = [i for i, (k, _) in enumerate(plan.aggregates()) if k == 12]
        assert len(hist_aggs) == 1
        plan.execute(ctx, whole)
        one = plan.analyzer_result(slot)
        # phase 1 alone leaves the aggregate pending and says so
        blobs = []
        for nm in names:
            plan.execute_partial(ctx, nm)
            blobs.append(plan.partial_export())
        plan.partial_reset()
        for b in blobs:
            plan.partial_merge(b)
        plan.finalize()
        if n > 4:
            assert plan.histogram_pending() == hist_aggs
            assert plan.analyzer_result(slot).error == 2
        _merged(ctx, plan, names)
        assert plan.histogram_pending() == []
        got = plan.analyzer_result(slot)
        assert got.error == 0 and got.u[0] == one.u[0]
        _hist_maps_equal(got.map, one.map)
        want = O.an_histogram(t, "v", nb)
        assert got.map["min"] == want["min"] and got.map["max"] == want["max"]
        for i, (lo, hi, cnt) in enumerate(want["buckets"]):
            assert got.map[f"bucket_{i}.lower"] == lo and got.map[f"bucket_{i}.upper"] == hi and got.map[f"bucket_{i}.count"] == cnt, i
    finally:
        for nm in names + [whole]:
            ctx.deregister_table(nm)


def test_histogram_shards_with_equal_ranges_need_no_second_phase(ctx):
    vals = np.tile(np.array([1.0, 2.0, 2.5, 9.0]), 1000)
    t = pa.table({"v": pa.array(vals)})
    names = []
    for i in range(3):
        names.append(f"mg_histeq_{i}")
        ctx.register_table(names[-1], t)
    try:
        plan = T.Plan()
        slot = T.HistogramAnalyzer("v", 4)._add_to(plan)
        blobs = []
        for nm in names:
            plan.execute_partial(ctx, nm)
            blobs.append(plan.partial_export())
        plan.partial_reset()
        for b in blobs:
            plan.partial_merge(b)
        plan.finalize()
        assert plan.histogram_pending() == []
        r = plan.analyzer_result(slot)
        assert r.error == 0 and [r.map[f"bucket_{i}.count"] for i in range(4)] == [9000.0, 0.0, 0.0, 3000.0]
    finally:
        for nm in names:
            ctx.deregister_table(nm)


def test_many_term_predicates_split_into_passes_instead_of_failing_the_suite(ctx):
    """12 three-term predicates (36 terms > SCAN_MAX_TERMS = 32) used to abort the whole plan; the reference
    evaluates every constraint on its own (core/suite.rs:84-100)"""
    rng = np.random.default_rng(3)
    n = 20_000
    cols = {f"c{i}": pa.array(rng.integers(-100, 100, n), mask=rng.random(n) < 0.05) for i in range(6)}
    t = pa.table(cols)
    ctx.register_table("mg_terms", t)
    try:
        exprs = [f"c{i % 6} >= {-50 + i} AND c{(i + 1) % 6} <= {60 - i} AND c{(i + 2) % 6} <> {i}" for i in range(14)]
        cb = T.Check.builder("many")
        for e in exprs:
            cb.satisfies(e)
        suite = T.ValidationSuite.builder("s").table_name("mg_terms").check(cb.build()).build()
        rs = suite.run(ctx).report.results
        assert len(rs) == len(exprs)
        for g, e in zip(rs, exprs):
            o = O.custom_sql(t, e)
            assert g.status.name.lower() == o.status and g.metric == o.metric and g.message == o.message, (e, g, o)
    finally:
        ctx.deregister_table("mg_terms")


# ---------------------------------------------------------------- the hand-written radix sort (K6 / K4) ----
def _gpu_sort(ctx, keys, begin_bit=0, n_passes=8):
    import ctypes as C
    from term_b200 import _ffi as F
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out_k = np.empty_like(keys)
    out_i = np.empty(len(keys), dtype=np.uint32)
    F.check(F.lib().tg_debug_sort_pairs(ctx.handle, keys.ctypes.data, len(keys), begin_bit, n_passes, out_k.ctypes.data, out_i.ctypes.data))
    return out_k, out_i


@pytest.mark.parametrize("n", [1, 2, 31, 5375, 5376, 5377, 100_003, 3_000_000])
@pytest.mark.parametrize("shape", ["random64", "small_range", "constant", "two_values", "f64_normal", "sorted", "reverse", "max_keys"])
def test_radix_sort_matches_stable_argsort(ctx, n, shape):
    """onesweep LSD sort: stable, every tile-boundary size, skewed digits (trivial passes are skipped), maximum keys (the
    value the padding of the last tile uses)"""
    rng = np.random.default_rng(n * 7 + len(shape))
    if shape == "random64":
        keys = rng.integers(0, 2**64, n, dtype=np.uint64)
    elif shape == "small_range":
        keys = rng.integers(0, 1000, n, dtype=np.uint64)
    elif shape == "constant":
        keys = np.full(n, 0x0123456789ABCDEF, dtype=np.uint64)
    elif shape == "two_values":
        keys = np.where(rng.random(n) < 0.5, np.uint64(5), np.uint64(2**63 + 5)).astype(np.uint64)
    elif shape == "f64_normal":
        b = rng.normal(100.0, 15.0, n).view(np.uint64)
        keys = np.where(b >> np.uint64(63), ~b, b | np.uint64(1 << 63))
    elif shape == "sorted":
        keys = np.sort(rng.integers(0, 2**40, n, dtype=np.uint64))
    elif shape == "reverse":
        keys = np.sort(rng.integers(0, 2**40, n, dtype=np.uint64))[::-1].copy()
    else:
        keys = np.where(rng.random(n) < 0.3, np.uint64(2**64 - 1), rng.integers(2**64 - 300, 2**64, n, dtype=np.uint64)).astype(np.uint64)
    got_k, got_i = _gpu_sort(ctx, keys)
    want_i = np.argsort(keys, kind="stable")
    assert (got_k == keys[want_i]).all()
    assert (got_i == want_i.astype(np.uint32)).all()


def test_radix_sort_on_a_bit_range_is_stable(ctx):
    rng = np.random.default_rng(77)
    keys = rng.integers(0, 2**64, 200_000, dtype=np.uint64)
    got_k, got_i = _gpu_sort(ctx, keys, begin_bit=32, n_passes=4)
    want_i = np.argsort(keys >> np.uint64(32), kind="stable")
    assert (got_i == want_i.astype(np.uint32)).all() and (got_k == keys[want_i]).all()
    got_k, got_i = _gpu_sort(ctx, keys, begin_bit=8, n_passes=2)
    want_i = np.argsort((keys >> np.uint64(8)) & np.uint64(0xFFFF), kind="stable")
    assert (got_i == want_i.astype(np.uint32)).all()


# ---------------------------------------------------------------- distributed Spearman: the device stages (tg_rank_*) ----
@pytest.mark.parametrize("n,world", [(10, 2), (5_000, 3), (400_000, 4)])
def test_rank_stages_sample_sort_emulated_on_one_gpu(built_lib, n, world):
    """the sample sort of term_b200.distributed.distributed_spearman with `world` engines on ONE device standing in for the
    ranks and device-to-device copies standing in for the NCCL all-to-all: global minimum ranks (ties, NULLs, an empty
    shard) must reproduce scipy's rankdata(method="min") correlation"""
    import torch
    from term_b200.distributed import GpuRankStages, choose_splitters
    rng = np.random.default_rng(n)
    x = np.round(rng.normal(0, 5, n), 0 if n > 100 else 1)
    y = np.round(0.6 * x + rng.normal(0, 3, n), 1)
    ids = rng.integers(-50, 50, n)  # an Int64 column as the second variable of a second pair
    t = pa.table({"x": pa.array(x, mask=rng.random(n) < 0.1), "y": pa.array(y, mask=rng.random(n) < 0.05), "k": pa.array(ids)})
    fr = [0.0, 0.3, 0.3, 0.7, 1.0][: world] + [1.0] if world > 2 else [0.0, 0.5, 1.0]
    cuts = [int(n * f) for f in fr]
    ctxs = [T.SessionContext(0) for _ in range(world)]
    try:
        for r, c in enumerate(ctxs):
            c.register_table("data", t.slice(cuts[r], cuts[r + 1] - cuts[r]))
        for cx, cy in (("x", "y"), ("x", "k")):
            st = [GpuRankStages(c) for c in ctxs]
            total = sum(s.begin("data", cx, cy) for s in st)

            def exchange():
                for s in st:
                    s.local_sort()
                samples = np.concatenate([s.sample(64) for s in st])
                sp = choose_splitters(samples, world)
                counts = [s.split(sp, world) for s in st]  # counts[src][dst]
                sends = [s.send() for s in st]
                n_recv = [sum(counts[src][dst] for src in range(world)) for dst in range(world)]
                recvs = [s.recv(n_recv[dst]) for dst, s in enumerate(st)]
                for dst in range(world):
                    o = 0
                    for src in range(world):
                        lo = sum(counts[src][:dst])
                        c = counts[src][dst]
                        if c:
                            recvs[dst][0][o: o + c].copy_(sends[src][0][lo: lo + c])
                            recvs[dst][1][o: o + c].copy_(sends[src][1][lo: lo + c])
                        o += c
                torch.cuda.synchronize()
                for dst, s in enumerate(st):
                    s.commit(n_recv[dst])
                return [sum(n_recv[:dst]) for dst in range(world)]

            bases = exchange()
            center = (total + 1.0) / 2.0
            for s, b in zip(st, bases):
                s.finish_x(b)
            bases = exchange()
            parts = [s.finish_y(b, center) for s, b in zip(st, bases)]
            assert sum(p[0] for p in parts) == total
            sx, sy, sxx, syy, sxy = (sum(p[1][k] for p in parts) for k in range(5))
            nn = float(total)
            num = nn * sxy - sx * sy
            den = math.sqrt((nn * sxx - sx * sx) * (nn * syy - sy * sy))
            rho = num / den if den else 0.0
            want = O.an_correlation(t, cx, cy, "spearman")
            assert abs(rho - want) <= 1e-9, (cx, cy, rho, want)
            # and the single-engine job (the same stages back to back) agrees
            ctxs[0].register_table("whole", t)
            one = T.CorrelationAnalyzer.spearman(cx, cy).compute(ctxs[0], "whole")
            ctxs[0].deregister_table("whole")
            assert one.u[0] == total and abs(one.metric_double - want) <= 1e-9
    finally:
        for c in ctxs:
            c.close()


# ---------------------------------------------------------------- regex features on the device ----
def test_word_boundaries_multiline_and_unicode_properties_on_the_device(ctx):
    """has_pattern with \\b / (?m) / \\p{..} / class set operations: the device walks the DFA the host compiled; counts must
    equal the oracle's (Python `re` / `regex` with the crate's semantics) on a corpus with multi-byte text and newlines"""
    from tests.test_regex_features import HAYSTACKS
    rng = np.random.default_rng(8)
    vals = [HAYSTACKS[int(i)] for i in rng.integers(0, len(HAYSTACKS), 20_000)]
    for i in range(0, len(vals), 97):
        vals[i] = None
    t = pa.table({"s": pa.array(vals, type=pa.string())})
    ctx.register_table("rxf", t.to_batches(max_chunksize=3001))
    try:
        pats = [r"\bfoo\b", r"foo\B", r"(?m)^foo$", r"(?m)^\w+$", r"\p{Lu}\p{Ll}+", r"\p{Script=Han}", r"[a-z&&[^aeiou]]+\b", r"(?i)\bHELLO\b",
                r"\b\d+\b", r"(?x) \b café \b"]
        cb = T.Check.builder("rx")
        for p in pats:
            cb.validates_regex("s", p, 0.1)
        cb.has_format("s", T.FormatType.Regex, 0.1, T.FormatOptions(case_sensitive=False, null_is_valid=False), arg=r"\bÉ\b")
        rs = T.ValidationSuite.builder("s").table_name("rxf").check(cb.build()).build().run(ctx).report.results
        want = [O.format_constraint(t, "s", "Regex", 0.1, arg=p) for p in pats]
        want.append(O.format_constraint(t, "s", "Regex", 0.1, arg=r"\bÉ\b", case_sensitive=False, null_is_valid=False))
        for g, o, p in zip(rs, want, pats + ["icase É"]):
            assert g.status.name.lower() == o.status and g.metric == o.metric and g.message == o.message, (p, g, o)
    finally:
        ctx.deregister_table("rxf")
