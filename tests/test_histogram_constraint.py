"""HistogramConstraint (constraints/histogram.rs): value frequencies + a host closure over the Histogram.
The reference's own tests (histogram.rs:440-769) pin the oracle on CPU and the product on the GPU; random columns compare
the product's buckets / entropy / message with the oracle's."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import term_oracle as O

# (values, assertion, description, expected status) — transcribed from the reference's tests, file:line in the ids
REF_CASES = [
    pytest.param(["A"] * 6 + ["B"] * 2 + ["C"] * 2, lambda h: h.most_common_ratio() < 0.5, "most common value appears less than 50%", "failure",
                 id="histogram.rs:517-546 most_common<0.5"),
    pytest.param(["A"] * 6 + ["B"] * 2 + ["C"] * 2, lambda h: h.most_common_ratio() < 0.7, None, "success", id="histogram.rs:548-556 most_common<0.7"),
    pytest.param(["RED", "BLUE", "GREEN", "YELLOW", "RED", "BLUE"], lambda h: 3 <= h.bucket_count() <= 5, "has between 3 and 5 distinct values", "success",
                 id="histogram.rs:558-581 bucket_count"),
    pytest.param(["A", "A", "B", "B", "C", "C", "D", "D"], lambda h: h.is_roughly_uniform(1.5), None, "success", id="histogram.rs:583-604 uniform"),
    pytest.param(["Popular1"] * 4 + ["Popular2"] * 3 + ["Rare1", "Rare2", "Rare3"], lambda h: h.follows_power_law(2, 0.7),
                 "top 2 values account for 70% of distribution", "success", id="histogram.rs:606-633 power_law"),
    pytest.param(["A", "A", None, None, None, "B", "B", "C"], lambda h: 0.3 < h.null_ratio() < 0.4, None, "success", id="histogram.rs:635-659 nulls"),
    pytest.param([], lambda h: True, None, "skipped", id="histogram.rs:661-671 empty"),
    pytest.param(["PENDING"] * 2 + ["APPROVED"] * 3 + ["REJECTED"], lambda h: (h.get_value_ratio("APPROVED") or 0.0) > 0.4, "APPROVED status is most common",
                 "success", id="histogram.rs:673-698 value_ratio"),
    pytest.param(["A"] * 4 + ["B"] * 3 + ["C"] * 2 + ["D"], lambda h: len(h.top_n(2)) == 2 and h.top_n(2)[0][1] == 0.4 and h.top_n(2)[1][1] == 0.3, None,
                 "success", id="histogram.rs:700-727 top_n"),
    pytest.param([25, 25, 30, 30, 30, 35, 35, 40, 45, 50], lambda h: h.bucket_count() >= 5 and h.most_common_ratio() < 0.4, "age distribution is reasonable",
                 "success", id="histogram.rs:729-768 int64"),
]


def _table(values):
    ints = any(isinstance(v, int) for v in values)
    return pa.table({"test_col": pa.array(values, type=pa.int64() if ints else pa.string())})


def test_histogram_struct_reference_vectors():
    """histogram.rs:442-515: the Histogram helpers on hand-built buckets"""
    h = O.OHistogram([("A", 50, 0.5), ("B", 30, 0.3), ("C", 20, 0.2)], 100, 0)
    assert (h.most_common_ratio(), h.least_common_ratio(), h.bucket_count(), h.null_ratio()) == (0.5, 0.2, 3, 0.0)
    uniform = O.OHistogram([(k, 25, 0.25) for k in "ABCD"], 100, 0)
    skewed = O.OHistogram([("A", 90, 0.9), ("B", 10, 0.1)], 100, 0)
    assert uniform.entropy() > skewed.entropy()


@pytest.mark.parametrize("values,assertion,description,want", REF_CASES)
def test_oracle_reference_cases(values, assertion, description, want):
    r = O.histogram_constraint(_table(values), "test_col", assertion, description or "custom assertion")
    assert r.status == want
    assert (r.message is not None) == (want != "success")


def _adapt(assertion):
    """the reference closures above read tuples from top_n: the product's Histogram.top_n gives the same (value, ratio) pairs"""
    return assertion


@pytest.mark.gpu
@pytest.mark.parametrize("values,assertion,description,want", REF_CASES)
def test_gpu_reference_cases(ctx, values, assertion, description, want):
    import term_b200.api as T
    t = _table(values)
    ctx.register_table("data", t)
    try:
        c = T.HistogramConstraint("test_col", assertion) if description is None else T.HistogramConstraint.new_with_description("test_col", assertion, description)
        g = c.evaluate(ctx, "data")
        o = O.histogram_constraint(t, "test_col", assertion, description or "custom assertion")
        assert g.status.name.lower() == want == o.status
        assert g.metric == o.metric and g.message == o.message and g.name == "histogram"
    finally:
        ctx.deregister_table("data")


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["utf8", "int64", "bool", "utf8_many"])
def test_gpu_histogram_matches_oracle_on_random_columns(ctx, kind):
    """buckets (value, count, ratio, order), counters, entropy and the failure message against the oracle; values whose
    VARCHAR order differs from their numeric order (-5 < 10 numerically, '-5' < '10' bytewise too, but '9' > '10')"""
    import term_b200.api as T
    rng = np.random.default_rng(5)
    n = 50_000
    mask = rng.random(n) < 0.07
    if kind == "utf8":
        vals = np.array(["alpha", "beta", "gamma", "", "δέλτα", "a much longer category name than eight bytes", "NULL"], dtype=object)[rng.integers(0, 7, n)]
        arr = pa.array(vals, type=pa.string(), mask=mask)
    elif kind == "utf8_many":
        vals = np.array([f"k{int(x):05d}" for x in rng.zipf(1.3, n) % 3000], dtype=object)
        arr = pa.array(vals, type=pa.string(), mask=mask)
    elif kind == "int64":
        arr = pa.array(rng.integers(-12, 13, n) * 1, mask=mask)
    else:
        arr = pa.array(rng.random(n) < 0.3, mask=mask)
    t = pa.table({"c": arr, "x": pa.array(rng.random(n))})
    ctx.register_table("hist_rand", t.to_batches(max_chunksize=7001))
    try:
        cons = T.HistogramConstraint.new_with_description("c", lambda h: h.most_common_ratio() < 0.01, "no value above 1%")
        plan = T.Plan()
        slot = cons._add_to(plan)
        plan.execute(ctx, "hist_rand")
        h = cons.histogram(plan, slot)
        oh = O.histogram_of(t, "c")
        assert [(b.value, b.count, b.ratio) for b in h.buckets] == oh.buckets
        assert (h.total_count, h.null_count, h.distinct_count) == (oh.total_count, oh.null_count, oh.distinct_count)
        g = cons._result(plan, slot)
        o = O.histogram_constraint(t, "c", lambda hh: hh.most_common_ratio() < 0.01, "no value above 1%")
        assert g.status.name.lower() == o.status == "failure" and g.metric == o.metric and g.message == o.message
        assert h.entropy() == oh.entropy() == g.metric
        # through a suite, next to other constraints of the same plan
        rs = (T.ValidationSuite.builder("s").table_name("hist_rand")
              .check(T.Check.builder("c").has_histogram("c", lambda hh: hh.bucket_count() == oh.distinct_count).completeness(["x"])
                     .has_histogram_with_description("c", lambda hh: hh.null_ratio() == 0.0, "no nulls").build()).build().run(ctx).report.results)
        assert [r.status.name for r in rs] == ["Success", "Success", "Failure"]
        assert rs[2].message.startswith("Histogram assertion 'no nulls' failed for column 'c'. Distribution: ")
        # a floating-point column is refused (Arrow's CAST(f64 AS VARCHAR) formatting is not restated), not mis-rendered
        bad = T.HistogramConstraint("x", lambda hh: True).evaluate(ctx, "hist_rand")
        assert bad.status.name == "Failure" and "not restated" in bad.message
        # an all-NULL column has nothing to analyze
    finally:
        ctx.deregister_table("hist_rand")


@pytest.mark.gpu
def test_gpu_histogram_all_null_column_is_skipped(ctx):
    import term_b200.api as T
    t = pa.table({"c": pa.array([None] * 1000, type=pa.string())})
    ctx.register_table("hist_null", t)
    try:
        g = T.HistogramConstraint("c", lambda h: True).evaluate(ctx, "hist_null")
        assert g.status.name == "Skipped" and g.message == "No data to analyze" and g.metric is None
    finally:
        ctx.deregister_table("hist_null")


@pytest.mark.parametrize("kind", ["utf8", "int64", "bool"])
def test_oracle_histogram_matches_arrow_value_counts(kind):
    """the oracle's buckets against Arrow's value_counts kernel (counts), with the reference query's order (count DESC, value ASC)
    and ratio (count / non-NULL rows) checked independently"""
    import pyarrow.compute as pc
    rng = np.random.default_rng(9)
    n = 20_000
    mask = rng.random(n) < 0.1
    if kind == "utf8":
        arr = pa.array(np.array(["x", "yy", "", "Zeta", "émile", "zz top"], dtype=object)[rng.zipf(1.5, n) % 6], type=pa.string(), mask=mask)
    elif kind == "int64":
        arr = pa.array(rng.integers(-20, 20, n), mask=mask)
    else:
        arr = pa.array(rng.random(n) < 0.7, mask=mask)
    t = pa.table({"c": arr})
    h = O.histogram_of(t, "c")
    vc = {}
    for e in pc.value_counts(arr).to_pylist():
        if e["values"] is None:
            continue
        v = e["values"]
        key = ("true" if v else "false") if kind == "bool" else str(v)
        vc[key] = e["counts"]
    assert {v: c for v, c, _ in h.buckets} == vc
    assert h.total_count == n and h.null_count == int(mask.sum())
    keys = [(-c, v.encode()) for v, c, _ in h.buckets]
    assert keys == sorted(keys)
    assert all(r == c / (n - h.null_count) for _, c, r in h.buckets)
    assert abs(sum(r for _, _, r in h.buckets) - 1.0) < 1e-12


def _grouped_partial(plan, values):
    """partial blob of one row shard of a value-histogram plan, built on the CPU: the grouped state is [u64 n] then per group
    {u32 key_len, key, u64 rows, u64 non-NULL rows} (term_b200/csrc/plan.cpp: grouped blob; the NULL group has 0 non-NULL rows
    and, in a value histogram, a key that no string value can equal — the column below holds the string 'NULL' too)"""
    import struct
    from collections import Counter
    (kind, _key), = plan.aggregates()
    counts = Counter(b"\xffNULL" if v is None else str(v).encode() for v in values)  # the NULL group's key is not valid UTF-8
    blob = struct.pack("<Q", len(counts))
    for kb, c in counts.items():
        blob += struct.pack("<I", len(kb)) + kb + struct.pack("<QQ", c, 0 if kb == b"\xffNULL" else c)
    pad = (-len(blob)) % 8
    rec = struct.pack("<QQ", kind, 0) + struct.pack("<8Q", len(values), 0, 0, 0, 0, 0, 0, 0) + struct.pack("<8d", *([0.0] * 8))
    rec += struct.pack("<Q", len(blob)) + blob + bytes(pad) + struct.pack("<Q", 0)
    return struct.pack("<Q", 1) + rec


@pytest.mark.parametrize("n_shards", [1, 3])
def test_value_histogram_finalize_and_shard_merge_on_cpu(built_lib, n_shards):
    """the host half of the histogram constraint without a GPU: shard states (here built by hand in the documented layout) merge
    through tg_plan_partial_merge, finalize orders the buckets (count DESC, value ASC), computes the entropy and leaves the
    assertion to the host closure — equal to the oracle on the whole column"""
    import term_b200.api as T
    rng = np.random.default_rng(2)
    vals = [None if rng.random() < 0.1 else ["b", "a", "cc", "", "NULL"][int(rng.integers(0, 5))] for _ in range(3000)]
    t = pa.table({"c": pa.array(vals, type=pa.string())})
    cons = T.HistogramConstraint.new_with_description("c", lambda h: h.bucket_count() == 4, "four categories")
    plan = T.Plan()
    slot = cons._add_to(plan)
    plan.partial_reset()
    for s in range(n_shards):
        plan.partial_merge(_grouped_partial(plan, vals[len(vals) * s // n_shards: len(vals) * (s + 1) // n_shards]))
    plan.finalize()
    h, oh = cons.histogram(plan, slot), O.histogram_of(t, "c")
    assert [(b.value, b.count, b.ratio) for b in h.buckets] == oh.buckets
    assert (h.total_count, h.null_count, h.distinct_count) == (oh.total_count, oh.null_count, 5)
    g = cons._result(plan, slot)
    o = O.histogram_constraint(t, "c", lambda hh: hh.bucket_count() == 4, "four categories")
    assert (g.status.name.lower(), g.metric, g.message) == (o.status, o.metric, o.message) and g.status.name == "Failure"
