"""Full-size parity through size-independent properties (BASELINE.json configs C2..C5 at the per-GPU shard size).

The oracle is a per-row Python/numpy restatement and takes minutes at 10^8 rows, so at these sizes the CUDA path is
checked against properties that hold at any size and against ground truth the device data was CONSTRUCTED to have,
recomputed with plain torch tensor ops on the same buffers (an implementation independent of every kernel under test):

* counts / completeness / min / max / Int64 sums / predicate counts / match counts / distinct counts: bit-exact;
* f64 mean 1e-9, stddev / correlation 1e-6 (the tolerances BASELINE.json `north_star` states);
* linearity: two half-table partials merged through tg_plan_partial_merge equal the one-pass answer;
* KLL quantiles within the rank-error bound of the reference's own accuracy harness.
"""
import math
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

REL_SUM, REL_MOMENT = 1e-9, 1e-6
N_C2 = 100_000_000   # BASELINE.json configs[1]
N_C3 = 25_000_000    # configs[2] / 8 GPUs
N_C4 = 125_000_000   # configs[3] / 8 GPUs
N_C5 = 125_000_000   # configs[4] / 8 GPUs


def _unpack(bits, n):
    import torch
    sh = torch.arange(8, device=bits.device, dtype=torch.uint8)
    nb = (n + 7) // 8
    return ((bits[:nb].unsqueeze(1) >> sh) & 1).bool().view(-1)[:n]


def _rel(a, b):
    return abs(a - b) / max(1.0, abs(b))


@pytest.fixture(scope="module")
def c2(ctx):
    import torch
    import bench
    dev = torch.device("cuda", 0)
    cols, keep = bench.make_device_table(torch, N_C2, 42 + 2, dev)
    ctx.register_device_table("c2_full", {k: {kk: vv for kk, vv in d.items() if kk not in ("tensor", "bits")} for k, d in cols.items()},
                              keepalive=keep)
    yield cols
    ctx.deregister_table("c2_full")
    del cols, keep
    torch.cuda.empty_cache()


def test_c2_full_numeric_set_against_torch_ground_truth(ctx, c2):
    import torch
    import bench
    import term_b200 as T
    n = N_C2
    suite = bench.build_full_suite(T, "c2_full")
    plan, slots = suite.build_plan()
    plan.execute(ctx, "c2_full")
    rs = [plan.result(s) for _, _, s in slots]
    # build_full_suite order: size, then per column completeness / min / max / mean / sum / stddev, then 4 correlations
    names = [f"f{k}" for k in range(4)] + [f"i{k}" for k in range(4)]
    assert len(rs) == 1 + 6 * len(names) + 4 and rs[0].name == "size" and rs[0].metric == float(n)
    masks = {}
    for ci, name in enumerate(names):
        r_comp, r_min, r_max, r_mean, r_sum, r_sd = rs[1 + 6 * ci: 7 + 6 * ci]
        assert [r.name for r in (r_comp, r_min, r_max, r_mean, r_sum, r_sd)] == ["completeness", "min", "max", "mean", "sum", "standard_deviation"]
        d = c2[name]
        v, m = d["tensor"][:n], _unpack(d["bits"], n)
        masks[name] = m
        cnt = int(m.sum().item())
        assert r_comp.metric == cnt / n, name
        if name[0] == "f":
            lo = torch.where(m, v, torch.full_like(v, math.inf)).min().item()
            hi = torch.where(m, v, torch.full_like(v, -math.inf)).max().item()
        else:
            lo = float(torch.where(m, v, torch.full_like(v, 2**62)).min().item())
            hi = float(torch.where(m, v, torch.full_like(v, -2**62)).max().item())
        assert r_min.metric == lo and r_max.metric == hi, name
        vv = v[m]
        if name[0] == "f":
            assert _rel(r_sum.metric, vv.sum().item()) <= REL_SUM, name
        else:
            assert r_sum.metric == float(vv.sum().item()), name  # exact i64 SUM (statistics.rs:278-308)
        vd = vv.to(torch.float64)
        assert _rel(r_mean.metric, vd.mean().item()) <= REL_SUM, name
        assert _rel(r_sd.metric, vd.std(unbiased=True).item()) <= REL_MOMENT, name
        del vv, vd
    pairs = (("f0", "f1"), ("f2", "f3"), ("i0", "i1"), ("f0", "i2"))
    for (a, b), r in zip(pairs, rs[-4:]):
        assert r.name == "correlation"
        m = masks[a] & masks[b]
        x, y = c2[a]["tensor"][:n][m].to(torch.float64), c2[b]["tensor"][:n][m].to(torch.float64)
        x = x - x.mean()
        y = y - y.mean()
        rho = ((x * y).sum() / torch.sqrt((x * x).sum() * (y * y).sum())).item()
        assert abs(r.metric - rho) <= REL_MOMENT, (a, b, r.metric, rho)
        del x, y, m


def test_c2_business_rules_and_predicate_counts(ctx, c2):
    import torch
    import bench
    import term_b200 as T
    n = N_C2
    suite = bench.build_suite(T, "c2_full")
    out = suite.run(ctx)
    rs = {r.name: r for r in out.report.results}
    f2, i0 = c2["f2"]["tensor"][:n], c2["i0"]["tensor"][:n]
    ok = (f2 > 0) & (i0 < 1_000_000) & _unpack(c2["f2"]["bits"], n) & _unpack(c2["i0"]["bits"], n)
    sat = int(ok.sum().item())
    assert rs["custom_sql"].metric == sat / n  # rows where the predicate is NULL are unsatisfied (custom_sql.rs:203-209)
    assert rs["custom_sql"].status == T.ConstraintStatus.Failure
    assert rs["size"].metric == float(n) and out.report.metrics.total_checks == 5


def test_c2_half_table_partials_merge_to_one_pass_answer(ctx, c2):
    """linearity: the row-sharded multi-GPU path in one process — two partial executions over disjoint halves merged in
    rank order; counts / min / max / i64 sums identical, moments within tolerance"""
    import bench
    import term_b200 as T
    n = N_C2
    h = (n // 2) // 64 * 64
    for nm, (start, rows) in (("c2_lo", (0, h)), ("c2_hi", (h, n - h))):
        ctx.register_device_table(nm, {k: dict(dtype=d["dtype"], n_rows=rows, values=d["values"] + 8 * start,
                                               validity=d["validity"] + start // 8) for k, d in c2.items()})
    try:
        suite = bench.build_full_suite(T, "c2_full")
        plan, slots = suite.build_plan()
        plan.execute(ctx, "c2_full")
        want = [plan.result(s) for _, _, s in slots]
        blobs = []
        for nm in ("c2_lo", "c2_hi"):
            plan.execute_partial(ctx, nm)
            blobs.append(plan.partial_export())
        plan.partial_reset()
        for b in blobs:
            plan.partial_merge(b)
        plan.finalize()
        got = [plan.result(s) for _, _, s in slots]
        for g, w in zip(got, want):
            assert g.status == w.status and g.name == w.name
            if g.name in ("size", "completeness", "min", "max"):
                assert g.metric == w.metric, (g, w)
            elif g.name == "sum":
                assert g.metric == w.metric or _rel(g.metric, w.metric) <= REL_SUM, (g, w)
            elif g.name == "mean":
                assert _rel(g.metric, w.metric) <= REL_SUM, (g, w)
            else:
                assert _rel(g.metric, w.metric) <= REL_MOMENT, (g, w)
    finally:
        ctx.deregister_table("c2_lo")
        ctx.deregister_table("c2_hi")


def test_c3_pattern_counts_against_byte_level_ground_truth(ctx):
    """25 M strings (avg 24 B): rows containing '@' are counted independently from the raw bytes (cumulative sum of
    the byte mask, differenced at the offsets); the email / SSN / card shapes are known by construction."""
    import torch
    import term_b200 as T
    from term_b200 import _ffi as F
    from tools.bench_suites import make_strings
    dev = torch.device("cuda", 0)
    n = N_C3
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 3)
    offs, data, v, (is_email, is_ssn, is_card), total = make_strings(n, g, dev)
    ctx.register_device_table("c3_full", {"s": dict(dtype=F.TG_UTF8, n_rows=n, values=data.data_ptr(), offsets=offs.data_ptr(),
                                                    validity=v.data_ptr(), n_value_bytes=total)}, keepalive=[offs, data, v])
    try:
        check = (T.Check.builder("pii").validates_regex("s", "@", 0.5).validates_email("s", 0.5).contains_ssn("s", 0.05)
                 .validates_credit_card("s", 0.5, True).has_min_length("s", 12).is_not_empty("s").build())
        rs = T.ValidationSuite.builder("c3").table_name("c3_full").check(check).build().run(ctx).report.results
        valid = _unpack(v, n)
        n_null = n - int(valid.sum().item())
        cs = torch.zeros(total + 1, dtype=torch.int32, device=dev)
        torch.cumsum((data[:total] == 64).to(torch.int32), 0, out=cs[1:])
        o = offs.to(torch.int64)
        has_at = (cs[o[1:]] - cs[o[:-1]]) > 0
        del cs
        lens = o[1:] - o[:-1]
        # NULL rows count as matching (`OR c IS NULL`, FormatOptions::null_is_valid default, format.rs:762-776;
        # length.rs:167-170), the denominator is COUNT(*)
        n_at = int((has_at & valid).sum().item())
        assert rs[0].metric == (n_at + n_null) / n
        assert rs[4].metric == (int(((lens >= 12) & valid).sum().item()) + n_null) / n
        assert rs[5].metric == (int(((lens >= 1) & valid).sum().item()) + n_null) / n
        # by construction: an email-shaped row is letters, '@' at len/2 and '.' four bytes before the end; it is a valid
        # address iff the first domain label is not empty, i.e. len - len/2 >= 6 (the '.' overwrites the '@' at len 8)
        n_email = int((is_email & (lens - lens // 2 >= 6) & valid).sum().item())
        assert rs[1].metric == (n_email + n_null) / n, (rs[1].metric, (n_email + n_null) / n)
        # SSN rows were built valid (area 1xx-5xx, non-zero group / serial), 11 bytes; card rows are 16 digits
        n_ssn = int((is_ssn & valid).sum().item())
        assert rs[2].metric == (n_ssn + n_null) / n, (rs[2].metric, (n_ssn + n_null) / n)
        n_card = int((is_card & valid).sum().item())
        assert rs[3].metric == (n_card + n_null) / n, (rs[3].metric, (n_card + n_null) / n)
    finally:
        ctx.deregister_table("c3_full")
        del offs, data, v
        torch.cuda.empty_cache()


def test_c4_uniqueness_and_foreign_key_counts_by_construction(ctx):
    """125 M keys: a permutation with d injected duplicates has exactly n - d distinct values; child keys drawn from
    [0, 1.0001 m) against parent 0..m-1 violate exactly where child >= m."""
    import torch
    import term_b200 as T
    from term_b200 import _ffi as F
    dev = torch.device("cuda", 0)
    n = N_C4
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 4)
    d = 1000
    keys = torch.zeros(n + 64, dtype=torch.int64, device=dev)
    keys[:n] = torch.randperm(n, generator=g, device=dev, dtype=torch.int64)
    keys[:d] = keys[d:2 * d]  # d values now appear twice, their d partners vanished
    keys[:n] *= 3  # sparse enough that it is not an id range
    ctx.register_device_table("c4_keys", {"k": dict(dtype=F.TG_INT64, n_rows=n, values=keys.data_ptr(), validity=None)}, keepalive=[keys])
    try:
        A = T.Assertion
        check = (T.Check.builder("u").validates_uniqueness(["k"], 0.9).validates_distinctness(["k"], A.GreaterThan(0.0))
                 .validates_unique_value_ratio(["k"], A.GreaterThan(0.0)).validates_primary_key(["k"]).build())
        rs = T.ValidationSuite.builder("c4u").table_name("c4_keys").check(check).build().run(ctx).report.results
        assert rs[0].metric == (n - d) / n and rs[1].metric == (n - d) / n
        assert rs[2].metric == (n - 2 * d) / n  # values occurring once / COUNT(*) (uniqueness.rs:656-690)
        assert rs[3].status == T.ConstraintStatus.Failure and rs[3].metric == d / n  # PK: (total - distinct) / total
    finally:
        ctx.deregister_table("c4_keys")
    del keys
    torch.cuda.empty_cache()
    m = n // 10
    parent = torch.zeros(m + 64, dtype=torch.int64, device=dev)
    parent[:m] = torch.randperm(m, generator=g, device=dev, dtype=torch.int64)
    child = torch.zeros(n + 64, dtype=torch.int64, device=dev)
    child[:n] = torch.randint(0, int(m * (1 + 1e-4)), (n,), generator=g, device=dev, dtype=torch.int64)
    from tools.bench_suites import validity
    cv = validity(n, g, dev, 0.01)
    ctx.register_device_table("c4_customers", {"id": dict(dtype=F.TG_INT64, n_rows=m, values=parent.data_ptr(), validity=None)}, keepalive=[parent])
    ctx.register_device_table("c4_orders", {"customer_id": dict(dtype=F.TG_INT64, n_rows=n, values=child.data_ptr(), validity=cv.data_ptr())},
                              keepalive=[child, cv])
    try:
        r = T.ForeignKeyConstraint("c4_orders.customer_id", "c4_customers.id").evaluate(ctx, "c4_orders")
        valid = _unpack(cv, n)
        bad = (child[:n] >= m) & valid
        n_null = n - int(valid.sum().item())
        # NULL child rows have no equal parent key: violations unless allow_nulls (foreign_key.rs:165-172)
        assert r.status == T.ConstraintStatus.Failure and r.metric == float(int(bad.sum().item()) + n_null)
        r2 = T.ForeignKeyConstraint("c4_orders.customer_id", "c4_customers.id").allow_nulls(True).evaluate(ctx, "c4_orders")
        assert r2.metric == float(int(bad.sum().item()))
    finally:
        ctx.deregister_table("c4_orders")
        ctx.deregister_table("c4_customers")
        del parent, child, cv
        torch.cuda.empty_cache()


def test_c5_kll_quantiles_within_rank_error_at_full_size(ctx):
    """KLL k=256 p50/p95/p99 on 125 M rows: rank error against the exact order statistics of the same column"""
    import torch
    import term_b200 as T
    from term_b200 import _ffi as F
    from tools.bench_suites import validity
    dev = torch.device("cuda", 0)
    n = N_C5
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 5)
    t = torch.zeros(n + 64, dtype=torch.float64, device=dev)
    t[:n].normal_(0.0, 1.0, generator=g).exp_()
    v = validity(n, g, dev, 0.05)
    ctx.register_device_table("c5_full", {"f": dict(dtype=F.TG_FLOAT64, n_rows=n, values=t.data_ptr(), validity=v.data_ptr())}, keepalive=[t, v])
    try:
        out = T.KllSketchAnalyzer("f", 256, (0.5, 0.95, 0.99)).compute(ctx, "c5_full")
        s, _ = torch.sort(t[:n][_unpack(v, n)])
        cnt = s.numel()
        assert out.map["count"] == float(cnt) and out.map["min"] == s[0].item() and out.map["max"] == s[-1].item()
        for q in (0.5, 0.95, 0.99):
            est = out.map[f"quantile_{q}"]
            lo = torch.searchsorted(s, torch.tensor([est], dtype=torch.float64, device=dev), right=False).item() / cnt
            hi = torch.searchsorted(s, torch.tensor([est], dtype=torch.float64, device=dev), right=True).item() / cnt
            err = 0.0 if lo <= q <= hi else min(abs(lo - q), abs(hi - q))
            assert err <= 0.01, (q, est, err)  # the reference harness' bound (tests/tpc_integration_tests.rs:533-551)
    finally:
        ctx.deregister_table("c5_full")
        del t, v
        torch.cuda.empty_cache()
