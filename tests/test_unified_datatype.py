"""The reference's second DataTypeConstraint (constraints/datatype.rs — what has_consistent_data_type builds): schema check,
the placeholder consistency check, and predicate validations counted over the non-NULL rows. The reference's own tests
(datatype.rs:474-622) pin the oracle on CPU and the product on the GPU."""
import numpy as np
import pyarrow as pa
import pytest

from oracle import term_oracle as O


def ref_tables():
    return {
        "specific": pa.table({"int_col": pa.array([1, 2, 3, 4, 5]), "string_col": pa.array(["a", "b", "c", "d", "e"])}),
        "non_negative": pa.table({"positive_values": pa.array([1.0, 2.0, 3.0, 0.0, 5.0]), "mixed_values": pa.array([1.0, -2.0, 3.0, 0.0, 5.0])}),
        "range": pa.table({"values": pa.array([10.0, 20.0, 30.0, 40.0, 50.0])}),
        "strings": pa.table({"strings": pa.array(["hello", "world", "", None, "test"])}),
    }


# (table, constructor name, args, expected status, expected metric or None) — datatype.rs file:line in the ids
REF_CASES = [
    pytest.param("specific", "specific_type", ("int_col", "Int64"), "success", 1.0, id="datatype.rs:491-495"),
    pytest.param("specific", "specific_type", ("int_col", "Utf8"), "failure", 0.0, id="datatype.rs:497-501"),
    pytest.param("non_negative", "non_negative", ("positive_values",), "success", 1.0, id="datatype.rs:534-538"),
    pytest.param("non_negative", "non_negative", ("mixed_values",), "failure", 0.8, id="datatype.rs:540-545"),
    pytest.param("range", "range", ("values", 0.0, 100.0), "success", 1.0, id="datatype.rs:570-582"),
    pytest.param("strings", "not_empty", ("strings",), "failure", 0.75, id="datatype.rs:607-621"),
]


def oracle_eval(t, ctor, args):
    import term_b200.api as T
    c = getattr(T.UnifiedDataTypeConstraint, ctor)(*args)  # (only its declared fields are read: the predicate text and description)
    actual = T.arrow_type_debug(t.schema.field(c.column).type)
    return O.unified_data_type(t, c.column, c.kind, c.predicate, c.description, c.threshold, c.expected, actual)


@pytest.mark.parametrize("table,ctor,args,status,metric", REF_CASES)
def test_oracle_reference_cases(table, ctor, args, status, metric, built_lib):
    r = oracle_eval(ref_tables()[table], ctor, args)
    assert r.status == status and r.metric == pytest.approx(metric, abs=1e-12)


def test_consistency_placeholder_and_messages(built_lib):
    import term_b200.api as T
    t = ref_tables()["range"]
    r = O.unified_data_type(t, "values", "consistency", threshold=0.9)
    assert (r.status, r.metric, r.message) == ("success", 0.95, "Type consistency 95.0% meets threshold 90.0%")
    r = O.unified_data_type(t, "values", "consistency", threshold=0.99)
    assert (r.status, r.metric, r.message) == ("failure", 0.95, "Type consistency 95.0% below threshold 99.0%")
    with pytest.raises(ValueError):
        T.UnifiedDataTypeConstraint.type_consistency("values", 1.5)
    with pytest.raises(ValueError):
        T.UnifiedDataTypeConstraint.custom("values", "{column} > 0; DROP TABLE x")
    assert T.arrow_type_debug(pa.timestamp("us")) == "Timestamp(Microsecond, None)" and T.arrow_type_debug(pa.timestamp("ms", "UTC")) == 'Timestamp(Millisecond, Some("UTC"))'
    assert T.arrow_type_debug(pa.float64()) == "Float64" and T.arrow_type_debug(pa.date32()) == "Date32" and T.arrow_type_debug(pa.large_string()) == "LargeUtf8"


@pytest.mark.gpu
@pytest.mark.parametrize("table,ctor,args,status,metric", REF_CASES)
def test_gpu_reference_cases(ctx, table, ctor, args, status, metric):
    import term_b200.api as T
    t = ref_tables()[table]
    ctx.register_table("data", t)
    try:
        g = getattr(T.UnifiedDataTypeConstraint, ctor)(*args).evaluate(ctx, "data")
        o = oracle_eval(t, ctor, args)
        assert g.status.name.lower() == status == o.status and g.metric == o.metric and g.message == o.message and g.name == "datatype"
    finally:
        ctx.deregister_table("data")


@pytest.mark.gpu
def test_gpu_unified_datatype_random_table(ctx, monkeypatch):
    import datetime as dt
    import term_b200.api as T
    monkeypatch.setenv("TG_FIXED_NOW_NS", str(int(dt.datetime(2024, 2, 10, 12, tzinfo=dt.timezone.utc).timestamp()) * 10**9))
    rng = np.random.default_rng(11)
    n = 30_000
    strs = np.array(["", "a", "héllo", "  ", "a longer string value"], dtype=object)[rng.integers(0, 5, n)]
    t = pa.table({
        "f": pa.array(rng.normal(1.0, 2.0, n), mask=rng.random(n) < 0.1), "i": pa.array(rng.integers(-3, 50, n), mask=rng.random(n) < 0.05),
        "whole": pa.array(np.where(rng.random(n) < 0.9, np.round(rng.normal(0, 10, n)), rng.normal(0, 10, n))),
        "s": pa.array(strs, type=pa.string(), mask=rng.random(n) < 0.2),
        "d": pa.array(rng.integers(19700, 19800, n).astype(np.int32), type=pa.date32(), mask=rng.random(n) < 0.1),
        "other": pa.array(rng.integers(0, 3, n), mask=rng.random(n) < 0.3)})
    ctx.register_table("udt", t.to_batches(max_chunksize=4000))
    U = T.UnifiedDataTypeConstraint
    cases = [("non_negative", ("f",)), ("positive", ("i",)), ("integer", ("i",)), ("range", ("f", -1.5, 4)), ("range", ("i", 0, 49)),
             ("not_empty", ("s",)), ("valid_utf8", ("s",)), ("max_bytes", ("s", 5)), ("past_date", ("d",)), ("future_date", ("d",)),
             ("date_range", ("d", "2024-01-01", "2024-03-01")), ("custom", ("i", "{column} + other >= 0")), ("custom", ("f", "other IS NULL OR other < 2")),
             ("specific_type", ("d", "Date32")), ("specific_type", ("s", "LargeUtf8")), ("type_consistency", ("i", 0.5)), ("type_consistency", ("i", 0.96))]
    try:
        cb = T.Check.builder("udt")
        for ctor, args in cases:
            cb.constraint(getattr(U, ctor)(*args))
        cb.has_consistent_data_type("f", 0.95)
        rs = T.ValidationSuite.builder("s").table_name("udt").check(cb.build()).build().run(ctx).report.results
        seen = set()
        for (ctor, args), g in zip(cases, rs):
            o = oracle_eval(t, ctor, args)
            assert g.status.name.lower() == o.status and g.metric == o.metric and g.message == o.message, (ctor, args, g, o)
            seen.add(g.metric)
        assert len(seen) > 8
        assert rs[-1].status.name == "Success" and rs[-1].message == "Type consistency 95.0% meets threshold 95.0%"
        # CAST(<float> AS INT) is outside the declared predicate grammar: an error result, not a wrong count
        unsupported = U.integer("whole").evaluate(ctx, "udt")
        assert unsupported.status.name == "Failure" and unsupported.metric is None
        missing = U.non_negative("nope").evaluate(ctx, "udt")
        assert missing.status.name == "Failure" and "nope" in missing.message
    finally:
        ctx.deregister_table("udt")
