"""CrossTableSumConstraint without grouping (constraints/cross_table_sum.rs): two sums on two tables compared on the host.
The reference's non-grouped tests (cross_table_sum.rs:733-783) pin the oracle on CPU and the product on the GPU."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import term_oracle as O


def tolerance_tables():  # cross_table_sum.rs:737-751
    return pa.table({"id": pa.array([1]), "total": pa.array([100.005])}), pa.table({"id": pa.array([1]), "amount": pa.array([100.001])})


def no_grouping_tables():  # cross_table_sum.rs:651-672 through :771-781
    return (pa.table({"id": pa.array([1, 2, 3, 4]), "customer_id": pa.array([1, 1, 2, 2]), "total": pa.array([100.0, 200.0, 150.0, 300.0])}),
            pa.table({"id": pa.array([1, 2]), "customer_id": pa.array([1, 2]), "amount": pa.array([300.0, 450.0])}))


def test_oracle_reference_cases():
    o, p = tolerance_tables()
    r = O.cross_table_sum(o, "total", p, "amount", "orders_tolerance.total", "payments_tolerance.amount")
    assert r.status == "failure" and r.metric == pytest.approx(0.004, abs=1e-9)
    assert r.message.startswith("Cross-table sum mismatch: 1/1 overall totals failed validation (exact match required). Examples: [Group 'ALL': "
                                "orders_tolerance.total = 100.0050, payments_tolerance.amount = 100.0010 (diff: 0.0040)]")
    r = O.cross_table_sum(o, "total", p, "amount", "orders_tolerance.total", "payments_tolerance.amount", tolerance=0.01)
    assert r.status == "success" and r.metric == pytest.approx(0.004, abs=1e-9)
    o, p = no_grouping_tables()
    r = O.cross_table_sum(o, "total", p, "amount", "orders_no_grouping.total", "payments_no_grouping.amount")
    assert (r.status, r.metric, r.message) == ("success", 0.0, None)


def test_qualified_names_and_configuration(built_lib):
    import term_b200.api as T
    C = T.CrossTableSumConstraint
    assert C.parse_qualified_column("orders.total") == ("orders", "total")  # cross_table_sum.rs:785-796
    for bad in ("invalid_column", "too.many.parts"):
        with pytest.raises(ValueError):
            C.parse_qualified_column(bad)
    c = C("orders.total", "payments.amount").group_by(["customer_id", "order_date"]).tolerance(-0.01).max_violations_reported(50)
    assert (c.group_by_columns, c._tolerance, c._max_violations) == (["customer_id", "order_date"], 0.01, 50)


@pytest.mark.gpu
def test_gpu_reference_cases_and_random_tables(ctx):
    import term_b200.api as T
    C = T.CrossTableSumConstraint
    o, p = tolerance_tables()
    ctx.register_table("orders_tolerance", o)
    ctx.register_table("payments_tolerance", p)
    o2, p2 = no_grouping_tables()
    ctx.register_table("orders_no_grouping", o2)
    ctx.register_table("payments_no_grouping", p2)
    rng = np.random.default_rng(17)
    n = 300_000
    # integer-valued doubles and Int64 cents: every summation order gives the same total, so the comparison is bit-exact
    big_l = pa.table({"v": pa.array(rng.integers(-10**6, 10**6, n).astype(np.float64), mask=rng.random(n) < 0.05), "c": pa.array(rng.integers(0, 10**4, n))})
    big_r = pa.table({"w": pa.array(rng.integers(-10**6, 10**6, n // 2).astype(np.float64)), "c": pa.array(rng.integers(0, 2 * 10**4, n // 2), mask=rng.random(n // 2) < 0.1),
                      "nulls": pa.array([None] * (n // 2), type=pa.float64())})
    ctx.register_table("big_l", big_l.to_batches(max_chunksize=50_000))
    ctx.register_table("big_r", big_r.to_batches(max_chunksize=50_000))
    try:
        cases = [(C("orders_tolerance.total", "payments_tolerance.amount"), (o, "total", p, "amount", 0.0, 100)),
                 (C("orders_tolerance.total", "payments_tolerance.amount").tolerance(0.01), (o, "total", p, "amount", 0.01, 100)),
                 (C("orders_no_grouping.total", "payments_no_grouping.amount"), (o2, "total", p2, "amount", 0.0, 100)),
                 (C("big_l.v", "big_r.w"), (big_l, "v", big_r, "w", 0.0, 100)),
                 (C("big_l.v", "big_r.w").tolerance(1e12), (big_l, "v", big_r, "w", 1e12, 100)),
                 (C("big_l.c", "big_r.c").max_violations_reported(0), (big_l, "c", big_r, "c", 0.0, 0)),
                 (C("big_l.c", "big_l.c"), (big_l, "c", big_l, "c", 0.0, 100)),
                 (C("big_l.v", "big_r.nulls").tolerance(5.0), (big_l, "v", big_r, "nulls", 5.0, 100))]
        for c, (lt, lc, rt, rc, tol, mv) in cases:
            g = c.evaluate(ctx)
            w = O.cross_table_sum(lt, lc, rt, rc, c.left_column, c.right_column, tol, mv)
            assert g.status.name.lower() == w.status and g.name == "cross_table_sum", (c.left_column, g, w)
            assert g.metric == pytest.approx(w.metric, rel=1e-12, abs=1e-9) and (g.message == w.message or lt is o), (c.left_column, g, w)
        rs = T.ValidationSuite.builder("s").table_name("big_l").check(
            T.Check.builder("c").cross_table_sum("orders_no_grouping.total", "payments_no_grouping.amount")
            .constraint(C("big_l.v", "big_r.w").group_by(["c"])).constraint(C("big_l.v", "nope.w")).build()).build().run(ctx).report.results
        assert [r.status.name for r in rs] == ["Success", "Failure", "Failure"] and "not supported" in rs[1].message
    finally:
        for name in ("orders_tolerance", "payments_tolerance", "orders_no_grouping", "payments_no_grouping", "big_l", "big_r"):
            ctx.deregister_table(name)
