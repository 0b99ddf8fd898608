"""JoinCoverageConstraint (constraints/join_coverage.rs) for single join keys whose probed side is unique: the unmatched rows come
from the foreign-key kernels. The reference's tests (join_coverage.rs:432-530) pin the oracle on CPU — which spells the joins out
for ANY key multiplicities — and the product on the GPU."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import term_oracle as O

REF = {  # (orders.customer_id, customers.id, expected rate) -> (status, metric)
    "success": ([1, 2, 3], [1, 2, 3], 1.0, "success", 1.0),          # join_coverage.rs:434-462
    "partial": ([1, 2, 999], [1, 2], 0.6, "success", 2 / 3),         # :467-496
    "failure": ([999, 998, 997], [1], 0.9, "failure", 0.0),          # :500-530
}


def tables(case):
    child, parent, *_ = REF[case]
    return pa.table({"id": pa.array(range(1, len(child) + 1)), "customer_id": pa.array(child)}), pa.table({"id": pa.array(parent)})


@pytest.mark.parametrize("case", list(REF))
def test_oracle_reference_cases(case):
    o, c = tables(case)
    _, _, rate, status, metric = REF[case]
    r = O.join_coverage(o, "customer_id", c, "id", "orders", "customers", rate)
    assert r.status == status and r.metric == pytest.approx(metric)
    if status == "failure":
        assert r.message == "Join coverage constraint failed: orders -> customers coverage is 0.00% (expected: 90.00%) (3 unmatched examples found)"


def test_oracle_join_multiplicities():
    left = pa.table({"k": pa.array([1, 1, 2, 3, None])})
    right = pa.table({"k": pa.array([1, 1, 2, 4, None])})
    # LEFT JOIN: 1,1 -> 2 rows each, 2 -> 1, 3 -> unmatched, NULL -> unmatched: 7 rows, 5 matched
    assert O.join_coverage(left, "k", right, "k", "l", "r", 0.0).metric == pytest.approx(5 / 7)
    assert O.join_coverage(left, "k", right, "k", "l", "r", 0.0, coverage="Right").metric == pytest.approx(5 / 7)
    assert O.join_coverage(left, "k", right, "k", "l", "r", 0.0, distinct_only=True).metric == pytest.approx(5 / 3)


@pytest.mark.gpu
def test_gpu_reference_cases_and_random_tables(ctx):
    import term_b200.api as T
    J, CT = T.JoinCoverageConstraint, T.CoverageType
    for case, (_, _, rate, status, metric) in REF.items():
        o, c = tables(case)
        ctx.register_table(f"orders_{case}", o)
        ctx.register_table(f"customers_{case}", c)
        try:
            g = J(f"orders_{case}", f"customers_{case}").on("customer_id", "id").expect_match_rate(rate).evaluate(ctx)
            w = O.join_coverage(o, "customer_id", c, "id", f"orders_{case}", f"customers_{case}", rate)
            assert g.status.name.lower() == status == w.status and g.metric == w.metric and g.message == w.message and g.name == "join_coverage"
        finally:
            ctx.deregister_table(f"orders_{case}")
            ctx.deregister_table(f"customers_{case}")
    rng = np.random.default_rng(23)
    n = 200_000
    dim = rng.permutation(150_000)[:100_000].astype(np.int64) * 7            # unique dimension keys
    fact = rng.integers(0, 150_000, n).astype(np.int64) * 7                  # ~2/3 of them find a partner
    uniq_fact = rng.permutation(150_000)[:90_000].astype(np.int64) * 7
    left = pa.table({"k": pa.array(fact, mask=rng.random(n) < 0.03)})
    right = pa.table({"k": pa.array(dim, mask=rng.random(len(dim)) < 0.02)})
    left_u = pa.table({"k": pa.array(uniq_fact, mask=rng.random(len(uniq_fact)) < 0.01)})
    empty = pa.table({"k": pa.array([], type=pa.int64())})
    for name, t in (("jc_left", left), ("jc_right", right), ("jc_left_u", left_u), ("jc_empty", empty)):
        ctx.register_table(name, t)
    try:
        cases = [(J("jc_left", "jc_right").on("k", "k").expect_match_rate(0.9), (left, right, "jc_left", "jc_right", 0.9, "Left", False, 100)),
                 (J("jc_left", "jc_right").on("k", "k").expect_match_rate(0.5), (left, right, "jc_left", "jc_right", 0.5, "Left", False, 100)),
                 (J("jc_left", "jc_right").on("k", "k").distinct_only(True).expect_match_rate(5.0), (left, right, "jc_left", "jc_right", 1.0, "Left", True, 100)),
                 (J("jc_left", "jc_right").on("k", "k").max_examples_reported(0), (left, right, "jc_left", "jc_right", 1.0, "Left", False, 0)),
                 (J("jc_left_u", "jc_right").on("k", "k").coverage_type(CT.RightCoverage).expect_match_rate(0.99), (left_u, right, "jc_left_u", "jc_right", 0.99, "Right", False, 100)),
                 (J("jc_left_u", "jc_right").on("k", "k").coverage_type(CT.BidirectionalCoverage).max_examples_reported(7), (left_u, right, "jc_left_u", "jc_right", 1.0, "Bidirectional", False, 7)),
                 (J("jc_empty", "jc_right").on("k", "k"), (empty, right, "jc_empty", "jc_right", 1.0, "Left", False, 100))]
        for c, (lt, rt, ln, rn, exp, cov, dist, mx) in cases:
            g = c.evaluate(ctx)
            w = O.join_coverage(lt, "k", rt, "k", ln, rn, exp, cov, dist, mx)
            same_metric = (g.metric != g.metric and w.metric != w.metric) or g.metric == w.metric
            assert g.status.name.lower() == w.status and same_metric and g.message == w.message, (cov, dist, g, w)
        # a duplicated key on the probed side multiplies the join's rows: refused, like composite keys and missing keys
        bad = [J("jc_right", "jc_left").on("k", "k"), J("jc_left", "jc_right").on_multiple([("k", "k"), ("k", "k")]), J("jc_left", "jc_right"),
               J("jc_left", "jc_right").on("k", "k").coverage_type(CT.RightCoverage)]
        rs = T.ValidationSuite.builder("s").table_name("jc_left").check(T.Check.builder("c").constraints(bad).join_coverage("jc_left", "jc_right").build()).build().run(ctx).report.results
        assert all(r.status.name == "Failure" and r.metric is None for r in rs)
        assert "not supported" in rs[0].message and "No join keys specified" in rs[2].message
    finally:
        for name in ("jc_left", "jc_right", "jc_left_u", "jc_empty"):
            ctx.deregister_table(name)
