"""TemporalOrderingConstraint (constraints/temporal_ordering.rs): before / after ordering, date ranges and business hours of the
timestamp columns of one table. The reference's tests (temporal_ordering.rs:607-668) pin the oracle on CPU and the product on
the GPU; random tables compare the product (the reference's comparisons as predicate counts) with the oracle (raw arrays)."""
import datetime as dt

import numpy as np
import pyarrow as pa
import pytest

from oracle import term_oracle as O


def ts(*texts, unit="us"):
    base = dt.datetime(1970, 1, 1)
    per = {"s": 1, "ms": 10**3, "us": 10**6, "ns": 10**9}[unit]
    out = []
    for t in texts:
        if t is None:
            out.append(None)
        else:
            d = dt.datetime.fromisoformat(t) - base
            out.append((d.days * 86400 + d.seconds) * per)
    return pa.array(out, type=pa.timestamp(unit))


def ordered():   # temporal_ordering.rs:611-626
    return pa.table({"id": pa.array([1, 2]), "created_at": ts("2024-01-01 10:00:00", "2024-01-01 11:00:00"),
                     "processed_at": ts("2024-01-01 10:05:00", "2024-01-01 11:10:00")})


def violated():  # temporal_ordering.rs:644-656
    return pa.table({"id": pa.array([1, 2]), "created_at": ts("2024-01-01 10:00:00", "2024-01-01 11:00:00"),
                     "processed_at": ts("2024-01-01 09:00:00", "2024-01-01 11:10:00")})


def test_oracle_reference_cases():
    r = O.temporal_ordering(ordered(), "before_after", ("created_at", "processed_at", False))
    assert (r.status, r.metric, r.message) == ("success", 1.0, None)
    r = O.temporal_ordering(violated(), "before_after", ("created_at", "processed_at", False))
    assert r.status == "failure" and r.metric == 0.5
    assert r.message == "Temporal ordering violation: 1 records where 'created_at' is not before 'processed_at' (50.00% compliance)"


def test_oracle_null_and_quirk_semantics():
    t = pa.table({"a": ts("2024-01-01 10:00:00", None, "2024-01-01 10:00:00", "2024-01-02 00:00:00"),
                  "b": ts("2024-01-01 10:00:00", "2024-01-01 10:00:00", None, "2024-01-01 00:00:00")})
    # NULL rows are filtered unless allowed through, where they count as violations; equal timestamps pass before_after (>=) and
    # fail before_or_equal (>): the reference's inverted flag
    assert O.temporal_ordering(t, "before_after", ("a", "b", False)).metric == 0.5
    assert O.temporal_ordering(t, "before_after", ("a", "b", True)).metric == 0.0
    assert O.temporal_ordering(t, "before_after", ("a", "b", False), allow_nulls=True).metric == 0.25
    assert O.temporal_ordering(pa.table({"a": ts(), "b": ts()}), "before_after", ("a", "b", False)).status == "success"
    # Monday 2024-01-01 09:00 inside, Saturday 2024-01-06 10:00 filtered by weekdays_only, 18:00 outside
    h = pa.table({"t": ts("2024-01-01 09:00:00", "2024-01-06 10:00:00", "2024-01-02 18:00:00", "2024-01-03 17:00:00", None)})
    assert O.temporal_ordering(h, "business_hours", ("t", "09:00", "17:00", False)).metric == 0.75
    assert O.temporal_ordering(h, "business_hours", ("t", "09:00", "17:00", True)).metric == pytest.approx(2 / 3)
    assert O.temporal_ordering(h, "date_range", ("t", "2024-01-02", "2024-01-04")).metric == 0.5


@pytest.mark.gpu
def test_gpu_reference_cases_and_random_tables(ctx):
    import term_b200.api as T
    C = T.TemporalOrderingConstraint
    for name, t, want in (("events_ordered", ordered(), "Success"), ("events_violated", violated(), "Failure")):
        ctx.register_table(name, t)
        try:
            g = C(name).before_after("created_at", "processed_at").evaluate(ctx)
            o = O.temporal_ordering(t, "before_after", ("created_at", "processed_at", False))
            assert g.status.name == want and g.status.name.lower() == o.status and g.metric == o.metric and g.message == o.message and g.name == "temporal_ordering"
        finally:
            ctx.deregister_table(name)
    rng = np.random.default_rng(21)
    n = 40_000
    for unit in ("s", "ms", "us", "ns"):
        per = {"s": 1, "ms": 10**3, "us": 10**6, "ns": 10**9}[unit]
        base = int(dt.datetime(2024, 1, 1, tzinfo=dt.timezone.utc).timestamp())
        a = (base + rng.integers(-40 * 86400, 40 * 86400, n)) * per + rng.integers(0, per, n)
        b = a + (rng.integers(-120, 600, n)) * per
        eq = rng.random(n) < 0.1
        b[eq] = a[eq]
        old = (rng.integers(-3 * 365 * 86400, 0, n)) * per   # before the epoch: negative values through the time-of-day arithmetic
        t = pa.table({"a": pa.array(a, type=pa.timestamp(unit), mask=rng.random(n) < 0.05), "b": pa.array(b, type=pa.timestamp(unit), mask=rng.random(n) < 0.05),
                      "old": pa.array(old, type=pa.timestamp(unit), mask=rng.random(n) < 0.05)})
        name = f"temporal_ord_{unit}"
        ctx.register_table(name, t.to_batches(max_chunksize=9000))
        try:
            cases = [(C(name).before_after("a", "b"), ("before_after", ("a", "b", False), False, 0)),
                     (C(name).before_or_equal("a", "b"), ("before_after", ("a", "b", True), False, 0)),
                     (C(name).before_after("a", "b").allow_nulls(True), ("before_after", ("a", "b", False), True, 0)),
                     (C(name).before_after("a", "b").tolerance_seconds(60), ("before_after", ("a", "b", False), False, 60)),
                     (C(name).before_or_equal("a", "b").tolerance_seconds(300).allow_nulls(True), ("before_after", ("a", "b", True), True, 300)),
                     (C(name).date_range("a", "2023-12-15", "2024-01-20 12:00:00"), ("date_range", ("a", "2023-12-15", "2024-01-20 12:00:00"), False, 0)),
                     (C(name).date_range("a", None, "2024-01-01").allow_nulls(True), ("date_range", ("a", None, "2024-01-01"), True, 0)),
                     (C(name).business_hours("a", "09:00", "17:00"), ("business_hours", ("a", "09:00", "17:00", False), False, 0)),
                     (C(name).business_hours("a", "08:30", "18:15").weekdays_only(True), ("business_hours", ("a", "08:30", "18:15", True), False, 0)),
                     (C(name).business_hours("old", "00:00", "11:59").weekdays_only(True).allow_nulls(True), ("business_hours", ("old", "00:00", "11:59", True), True, 0))]
            seen = set()
            for c, (kind, args, nulls, tol) in cases:
                g = c.evaluate(ctx)
                o = O.temporal_ordering(t, kind, args, nulls, tol)
                assert g.status.name.lower() == o.status and g.metric == o.metric and g.message == o.message, (unit, kind, args, nulls, tol, g, o)
                seen.add(g.metric)
            assert len(seen) >= 8
            # inside a suite on another table the constraint still reads its own; unsupported validations are error results
            rs = T.ValidationSuite.builder("s").table_name(name).check(
                T.Check.builder("c").constraint(C(name).before_after("a", "b")).constraint(C(name).max_time_gap("a", 60))
                .constraint(C(name).business_hours("a", "09:00", "17:00").with_timezone("Europe/Paris")).temporal_ordering(name).build()).build().run(ctx).report.results
            assert rs[0].metric == O.temporal_ordering(t, "before_after", ("a", "b", False)).metric
            assert [r.status.name for r in rs[1:]] == ["Failure"] * 3 and all(r.metric is None for r in rs[1:])
        finally:
            ctx.deregister_table(name)
