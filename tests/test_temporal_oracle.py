"""CPU checks of the oracle's constant date / timestamp arithmetic (now() / current_date / INTERVAL folding of `satisfies`):
day-time intervals against datetime.timedelta, calendar months against hand-checked month ends, and the exact-nanosecond
comparison of a coarser column against an instant between two of its ticks."""
import datetime as dt

import numpy as np
import pyarrow as pa
import pytest

from oracle import term_oracle as O

EPOCH = dt.datetime(1970, 1, 1, tzinfo=dt.timezone.utc)


def ns_of(d: dt.datetime) -> int:
    delta = d - EPOCH
    return (delta.days * 86400 + delta.seconds) * 10**9 + delta.microseconds * 1000


def test_interval_literals():
    assert O.interval_literal("1 day") == (0, 86400 * 10**9)
    assert O.interval_literal("2 hours 30 minutes") == (0, 9000 * 10**9)
    assert O.interval_literal("1.5 hours") == (0, 5400 * 10**9)
    assert O.interval_literal("1 year 2 months 3 days") == (14, 3 * 86400 * 10**9)
    assert O.interval_literal("-3 days") == (0, -3 * 86400 * 10**9)
    assert O.interval_literal("250 milliseconds 7 microseconds 9 nanoseconds") == (0, 250_007_009)
    assert O.interval_literal("2 WEEKS") == (0, 14 * 86400 * 10**9)
    with pytest.raises(Exception):
        O.interval_literal("soon")


@pytest.mark.parametrize("text,delta", [("1 day", dt.timedelta(days=1)), ("36 hours", dt.timedelta(hours=36)), ("90 minutes 15 seconds", dt.timedelta(minutes=90, seconds=15)),
                                        ("1 week 1 day", dt.timedelta(days=8))])
def test_day_time_intervals_match_timedelta(text, delta):
    base = dt.datetime(2024, 2, 28, 23, 59, 30, tzinfo=dt.timezone.utc)
    for sign in (1, -1):
        e = ("ar", "+" if sign > 0 else "-", ("lit", "2024-02-28 23:59:30"), ("interval", text))
        assert O.temporal_const_ns(e, 0) == ns_of(base + sign * delta)


@pytest.mark.parametrize("start,months,sign,want", [("2024-03-31", 1, -1, "2024-02-29"), ("2023-03-31", 1, -1, "2023-02-28"), ("2024-01-31", 1, 1, "2024-02-29"),
                                                    ("2024-11-30", 3, 1, "2025-02-28"), ("2024-01-15", 14, -1, "2022-11-15"), ("1969-12-31", 2, 1, "1970-02-28")])
def test_calendar_months_clamp_to_the_month_end(start, months, sign, want):
    e = ("ar", "+" if sign > 0 else "-", ("lit", start + " 10:00:00"), ("interval", f"{months} months"))
    assert O.temporal_const_ns(e, 0) == O.temporal_literal(want + " 10:00:00", "n")


def test_now_and_current_date(monkeypatch):
    now = ns_of(dt.datetime(2024, 1, 31, 10, 11, 12, 500000, tzinfo=dt.timezone.utc))
    monkeypatch.setenv("TG_FIXED_NOW_NS", str(now))
    assert O.query_now_ns() == now
    assert O.temporal_const_ns(("fn", "NOW", []), now) == now
    assert O.temporal_const_ns(("fn", "CURRENT_DATE", []), now) == O.temporal_literal("2024-01-31", "n")
    # Date32 against a time of day / seconds against half a second: no tick equals the instant, > and >= agree, < and <= agree
    d = np.array([19752, 19753, 19754], dtype=np.int32)  # 2024-01-30 .. 2024-02-01
    s = np.array([now // 10**9 - 1, now // 10**9, now // 10**9 + 1])
    t = pa.table({"d": pa.array(d, type=pa.date32()), "ts": pa.array(s, type=pa.timestamp("s"))})
    want = {"d > now()": 1, "d >= now()": 1, "d < now()": 2, "d <= now()": 2, "d = now()": 0, "d <> now()": 3, "d = current_date": 1,
            "ts > now()": 1, "ts >= now()": 1, "ts < now()": 2, "ts <= now()": 2, "ts = now()": 0, "now() > ts": 2,
            "ts > now() - interval '1 day'": 3, "ts > now() + interval '500 milliseconds'": 0, "ts >= now() + interval '500 milliseconds'": 1}
    for p, k in want.items():
        r = O.custom_sql(t, p)
        assert r.metric == k / 3, (p, r)


def test_month_arithmetic_matches_pandas_dateoffset():
    """calendar-month addition (day of month clamped to the target month's length) cross-checked against pandas' DateOffset on
    random instants, both directions, across leap years and the epoch"""
    pd = pytest.importorskip("pandas")
    rng = np.random.default_rng(3)
    for _ in range(400):
        secs = int(rng.integers(-2_000_000_000, 4_000_000_000))
        months = int(rng.integers(0, 40))
        sign = 1 if rng.random() < 0.5 else -1
        ts = pd.Timestamp(secs, unit="s", tz="UTC")
        want = ts + pd.DateOffset(months=sign * months)
        got = O._add_interval(secs * 10**9, months, 0, sign)
        assert got == int(want.value), (ts, months, sign)


def test_day_time_arithmetic_matches_arrow_compute():
    """timestamp +/- a day-time interval against Arrow's own timestamp + duration kernel"""
    import pyarrow.compute as pc
    rng = np.random.default_rng(4)
    base = rng.integers(-10**9, 2 * 10**9, 200)
    ts = pa.array(base * 10**9, type=pa.timestamp("ns"))
    for text, ns in (("1 day", 86400 * 10**9), ("90 minutes", 5400 * 10**9), ("2 weeks 3 hours", (14 * 86400 + 10800) * 10**9), ("250 milliseconds", 250 * 10**6)):
        months, nanos = O.interval_literal(text)
        assert (months, nanos) == (0, ns)
        want = pc.add(ts, pa.scalar(ns, type=pa.duration("ns"))).cast(pa.int64()).to_pylist()
        got = [O._add_interval(int(b) * 10**9, 0, nanos, 1) for b in base]
        assert got == want


def test_product_parser_accepts_the_temporal_grammar(built_lib):
    """parser level, no GPU: now() / CURRENT_TIMESTAMP / current_date, INTERVAL '..' and INTERVAL '..' UNIT parse at plan-build time
    (the folding against the column's unit happens at execute); malformed forms are SQL parser errors reported as failed
    constraints, like the reference's planning errors (custom_sql.rs:212-232)"""
    import term_b200.api as T

    def built(expr):
        plan = T.Plan()
        slot = T.CustomSqlConstraint(expr)._add_to(plan)
        plan.finalize()
        return plan.result(slot)

    for ok in ("ts > now() - interval '1 day'", "ts >= CURRENT_TIMESTAMP - INTERVAL '2' MONTH", "d = current_date", "d < today() + INTERVAL '1 week 2 days'",
               "TIMESTAMP '2024-01-01 00:00:00' + interval '90 minutes' <= ts", "ts BETWEEN now() - interval '1 hour' AND now()"):
        r = built(ok)
        assert r.status.name == "Skipped" and r.message == "No data to validate", (ok, r)   # parsed; nothing executed yet
    for bad in ("ts > now( - 1", "x > interval '1 day' +", "ts > now() - interval"):
        r = built(bad)
        assert (r.status.name == "Failure" and "SQL" in r.message) or bad.endswith("interval"), (bad, r)
