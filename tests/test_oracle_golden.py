"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(tests/golden/reference_vectors.json, each case citing its Rust test). CPU only."""
import math

import numpy as np
import pytest

from . import helpers as H

CASES = H.load_golden()
CONSTRAINT_CASES = [c for c in CASES if c["op"]["kind"] not in ("analyzer", "assertion", "logical")]
ANALYZER_CASES = [c for c in CASES if c["op"]["kind"] == "analyzer"]


@pytest.mark.parametrize("case", CONSTRAINT_CASES, ids=[c["id"] for c in CONSTRAINT_CASES])
def test_oracle_constraint_matches_reference_vector(case):
    r = H.oracle_eval(case)
    H.check_expect(r.status, r.metric, r.message, case["expect"], case["ref"])


@pytest.mark.parametrize("case", ANALYZER_CASES, ids=[c["id"] for c in ANALYZER_CASES])
def test_oracle_analyzer_matches_reference_vector(case):
    got = H.oracle_analyzer(case)
    exp = case["expect"]
    if exp.get("no_data"):
        assert got.get("no_data")
        return
    for key in ("u", "f"):
        if key in exp:
            assert list(got[key][: len(exp[key])]) == exp[key], (key, got, exp)
    if "metric" in exp:
        tol = exp.get("metric_tol", 0.0)
        assert abs(got["metric"] - exp["metric"]) <= tol, (got, exp)
    if "metric_long" in exp:
        assert got["metric_long"] == exp["metric_long"]
    if "metric_gt" in exp:
        assert exp["metric_gt"] < got["metric"] < exp["metric_lt"]
    if "map" in exp:
        assert got["map"] == exp["map"] and got["n_groups"] == exp["n_groups"]


def test_oracle_assertion_and_logical_vectors():
    from oracle import term_oracle as O
    for c in CASES:
        if c["op"]["kind"] == "assertion":
            for a, v, want in c["op"]["cases"]:
                assert O.assertion_eval(tuple(a), v) is want, (a, v)
            for a, want in c["op"]["descriptions"]:
                assert O.assertion_desc(tuple(a)) == want
        if c["op"]["kind"] == "logical":
            for op, vals, want in c["op"]["cases"]:
                assert O.logical_eval(tuple(op), vals) is want, (op, vals)


def test_rust_f64_display():
    from oracle.term_oracle import rust_f64
    assert rust_f64(20.0) == "20"
    assert rust_f64(0.1) == "0.1"
    assert rust_f64(1e21) == "1000000000000000000000"
    assert rust_f64(1e-7) == "0.0000001"
    assert rust_f64(-2.5) == "-2.5"
    assert rust_f64(float("nan")) == "NaN" and rust_f64(float("inf")) == "inf"


# ---- the reference's in-tree KllSketch restated (oracle/kll_ref.py) against its own unit tests ----
def test_kll_ref_basic_operations():  # kll_sketch.rs:406-425
    from oracle.kll_ref import KllSketch
    s = KllSketch(100)
    assert s.n == 0
    for i in range(1000):
        s.update(float(i))
    assert s.count() == 1000
    med = s.get_quantile(0.5)
    assert 0.0 <= med <= 999.0
    assert s.get_quantile(0.0) == 0.0 and s.get_quantile(1.0) == 999.0


def test_kll_ref_single_value_nan_and_merge():  # kll_sketch.rs:427-458, 520-558
    from oracle.kll_ref import KllSketch
    s = KllSketch(50)
    s.update(42.0)
    assert s.get_quantile(0.0) == 42.0 and s.get_quantile(0.5) == 42.0 and s.get_quantile(1.0) == 42.0
    s = KllSketch(50)
    s.update(1.0)
    s.update(float("nan"))
    s.update(2.0)
    assert s.count() == 2
    a, b = KllSketch(100), KllSketch(100)
    for i in range(500):
        a.update(float(i))
    for i in range(500, 1000):
        b.update(float(i))
    a.merge(b)
    assert a.count() == 1000
    assert a.get_quantile(0.0) == 0.0 and a.get_quantile(1.0) == 999.0
    assert abs(KllSketch(200).relative_error_bound() - 1.65 / math.sqrt(200)) < 1e-12


def test_kll_ref_compactor_halves():  # kll_sketch.rs:497-518: 4 items -> 2 kept + 2 promoted
    from oracle.kll_ref import Compactor
    c = Compactor(4)
    c.items = [1.0, 2.0, 3.0, 4.0]
    promoted = c.compact()
    assert len(promoted) == 2 and len(c.items) == 2
    assert sorted(promoted + c.items) == [1.0, 2.0, 3.0, 4.0]


def test_siphash13_known_answer():
    # SipHash-1-3 of the empty message with zero keys (Rust: DefaultHasher::new().finish())
    from oracle.kll_ref import siphash13
    assert siphash13([]) == 0xD1FBA762150C532C


# ---- oracle self-consistency on the property-test generator (tests/property_tests.rs:102-145) ----
def _property_column(n, null_fraction, lo, rng_range):
    nulls = round(n * null_fraction)
    vals = [None] * nulls + [lo + (j / max(1, n - nulls)) * rng_range for j in range(n - nulls)]
    return vals


@pytest.mark.parametrize("n,frac", [(10, 0.0), (100, 0.25), (1000, 0.5), (7, 1.0)])
def test_oracle_property_generators(n, frac):
    from oracle import term_oracle as O
    import pyarrow as pa
    vals = _property_column(n, frac, -50.0, 100.0)
    t = pa.table({"c": pa.array(vals, type=pa.float64())})
    nn = sum(v is not None for v in vals)
    r = O.completeness(t, "c", 0.0)
    assert r.metric == nn / n  # property_tests.rs:215-257
    r = O.size(t, ("Equals", float(n)))
    assert r.status == "success" and r.metric == float(n)  # :309-365
    if nn:
        arr = np.array([v for v in vals if v is not None])
        assert abs(O.stat_value(O.table_cols(t)["c"], "Mean") - arr.mean()) < 1e-9  # :378-425
        assert O.stat_value(O.table_cols(t)["c"], "Min") == arr.min()  # :430-480
        assert O.stat_value(O.table_cols(t)["c"], "Max") == arr.max()
    if nn > 1:
        assert abs(O.stat_value(O.table_cols(t)["c"], "StandardDeviation") - arr.std(ddof=1)) < 1e-9  # :776-825


# ---- second opinion on the oracle's SQL-aggregate semantics: Arrow's own compute kernels (the library DataFusion's
# aggregates are built on; SURVEY §8c "pyarrow.compute for cross-checks") on seeded random columns ----
@pytest.mark.parametrize("seed", range(6))
def test_oracle_aggregates_agree_with_arrow_compute(seed):
    import math
    import numpy as np
    import pyarrow as pa
    import pyarrow.compute as pc
    from oracle import term_oracle as O
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 4000))
    f = pa.array(np.round(rng.normal(10.0, 4.0, n), 2), mask=rng.random(n) < 0.2)
    i = pa.array(rng.integers(-40, 40, n), mask=rng.random(n) < 0.2)
    s = pa.array([f"k{v}" for v in rng.integers(0, 30, n)], mask=rng.random(n) < 0.2)
    t = pa.table({"f": f, "i": i, "s": s})
    for col in ("f", "i", "s"):
        tot, nn, _ = O.an_completeness(t, col)
        assert (tot, nn) == (n, pc.count(t.column(col)).as_py())  # COUNT(*), COUNT(c)
        assert O.distinct_counts(t, [col])["distinct_nonnull"] == pc.count_distinct(t.column(col), mode="only_valid").as_py()
    cols = O.table_cols(t)
    for col in ("f", "i"):
        arr = t.column(col)
        if pc.count(arr).as_py() == 0:
            assert O.stat_value(cols[col], "Min") is None
            continue
        mm = pc.min_max(arr).as_py()
        assert O.stat_value(cols[col], "Min") == float(mm["min"]) and O.stat_value(cols[col], "Max") == float(mm["max"])
        assert abs(O.stat_value(cols[col], "Mean") - pc.mean(arr).as_py()) <= 1e-9 * max(1.0, abs(pc.mean(arr).as_py()))
        assert abs(O.stat_value(cols[col], "Sum") - float(pc.sum(arr).as_py())) <= 1e-9 * max(1.0, abs(float(pc.sum(arr).as_py())))
        if pc.count(arr).as_py() >= 2:  # STDDEV / VARIANCE are the sample forms (ddof = 1)
            sd, var = pc.stddev(arr, ddof=1).as_py(), pc.variance(arr, ddof=1).as_py()
            assert abs(O.stat_value(cols[col], "StandardDeviation") - sd) <= 1e-9 * max(1.0, sd)
            assert abs(O.stat_value(cols[col], "Variance") - var) <= 1e-9 * max(1.0, var)
        else:
            assert O.stat_value(cols[col], "StandardDeviation") is None
    # COUNT(CASE WHEN p THEN 1 END): rows where the predicate is NULL do not count
    want = pc.sum(pc.and_kleene(pc.greater(t.column("f"), 10.0), pc.less(t.column("i"), 5)).cast(pa.int64())).as_py() or 0
    got = O.predicate_counts(t, "f > 10 AND i < 5")
    assert got[0] == want and math.isfinite(float(want))


# ---- string sub-expressions of `satisfies` predicates: Arrow's own string kernels as the second opinion (DataFusion's LIKE,
# character_length and Utf8 comparisons are these kernels' Rust siblings: `%` / `_` / backslash escape, characters not bytes,
# byte-wise order) ----
@pytest.mark.parametrize("seed", range(4))
def test_oracle_string_predicates_agree_with_arrow_compute(seed):
    import numpy as np
    import pyarrow as pa
    import pyarrow.compute as pc
    from oracle import term_oracle as O
    rng = np.random.default_rng(100 + seed)
    n = 600
    alphabet = list("ab_%c ") + ["é", "你", "🦀"]
    mk = lambda: ["".join(rng.choice(alphabet, rng.integers(0, 7))) for _ in range(n)]
    s = pa.array(mk(), type=pa.string(), mask=rng.random(n) < 0.15)
    u = pa.array(mk(), type=pa.string(), mask=rng.random(n) < 0.15)
    t = pa.table({"s": s, "u": u})

    def count_true(arr):
        return pc.sum(pc.fill_null(arr, False).cast(pa.int64())).as_py() or 0

    for pat in ["a%", "%a", "%a%b%", "_", "__%", "%", "", "a\\%%", "%\\_%", "_é%", "%🦀", "ab c", "%你_"]:
        want = count_true(pc.match_like(s, pat))
        assert O.predicate_counts(t, f"s LIKE '{pat}'") == (want, n), pat
        want_not = count_true(pc.invert(pc.match_like(s, pat)))
        assert O.predicate_counts(t, f"s NOT LIKE '{pat}'") == (want_not, n), pat
    for k in (0, 1, 3, 6):
        assert O.predicate_counts(t, f"LENGTH(s) >= {k}")[0] == count_true(pc.greater_equal(pc.utf8_length(s), k))
        assert O.predicate_counts(t, f"OCTET_LENGTH(s) = {k}")[0] == count_true(pc.equal(pc.binary_length(s), k))
    for op, fn in (("=", pc.equal), ("<>", pc.not_equal), ("<", pc.less), ("<=", pc.less_equal), (">", pc.greater), (">=", pc.greater_equal)):
        assert O.predicate_counts(t, f"s {op} u")[0] == count_true(fn(s.cast(pa.binary()), u.cast(pa.binary()))), op
        assert O.predicate_counts(t, f"s {op} 'b'")[0] == count_true(fn(s.cast(pa.binary()), pa.scalar(b"b"))), op
        assert O.predicate_counts(t, f"'b' {op} s")[0] == count_true(fn(pa.scalar(b"b"), s.cast(pa.binary()))), op


def test_oracle_temporal_literals_agree_with_arrow_casts():
    """string literals compared with date / timestamp columns: the oracle's cast against Arrow's own string -> Date32 / Timestamp casts"""
    import pyarrow as pa
    import pyarrow.compute as pc
    from oracle import term_oracle as O
    dates = ["1970-01-01", "2024-01-31", "2024-02-29", "1969-12-31", "1999-12-31", "2100-03-01", "0001-01-01"]
    want = pc.cast(pa.array(dates), pa.date32()).cast(pa.int32()).to_pylist()
    assert [O.temporal_literal(d, "D") for d in dates] == want
    stamps = ["2024-01-31", "2024-01-31 10:11:12", "2024-01-31T10:11:12", "2024-01-31 10:11:12.5", "2024-01-31T10:11:12.123456789",
              "1969-12-31 23:59:59.25", "2024-01-31T10:11:12Z", "2024-01-31T12:11:12+02:00", "2024-01-31 05:41:12.5-04:30", "2024-01-31 10:11"]
    for unit, u in (("s", "s"), ("ms", "m"), ("us", "u"), ("ns", "n")):
        for text in stamps:
            # (Arrow refuses to drop sub-unit digits on a safe cast: the literal path truncates like DataFusion's cast to the column type)
            got = O.temporal_literal(text, u)
            has_zone = text.endswith("Z") or "+" in text or "-" in text[11:]
            ref = pc.cast(pa.array([text]), pa.timestamp("ns", tz="UTC") if has_zone else pa.timestamp("ns"))
            ref_ns = ref.cast(pa.int64()).to_pylist()[0]
            div = {"s": 10**9, "m": 10**6, "u": 10**3, "n": 1}[u]
            assert got == ref_ns // div, (text, unit, got, ref_ns)
    for bad in ("2024-13-01", "2024-02-30", "31/01/2024", "2024-01-31 25:00:00", "abc", ""):
        for u in ("D", "u"):
            with pytest.raises(ValueError):
                O.temporal_literal(bad, u)
