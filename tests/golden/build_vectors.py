#!/usr/bin/env python3
"""Writes tests/golden/reference_vectors.json: the known-answer tests the reference holds for the hot
path, transcribed by hand from its Rust test modules (the reference is Rust and cannot run here, so
these vectors are inputs + asserted outputs copied from the cited tests, not outputs of a run).

Each case: id, ref (file:line under /root/reference/term-guard/src unless noted), tables, op, expect.
Column encoding: {"type": "i64"|"f64"|"str", "values": [... null ...]}.
expect: status, metric (exact unless metric_tol / metric_gt / metric_lt), message_contains [..].
"""
import json
import os


def col(t, v):
    return {"type": t, "values": v}


def tbl(**cols):
    return {"data": cols}


C = []


def case(id, ref, tables, op, **expect):
    C.append({"id": id, "ref": ref, "tables": tables, "op": op, "expect": expect})


# ------------------------------------------------------------------ completeness ----
case("completeness_single_complete", "constraints/completeness.rs:339-352",
     tbl(id=col("i64", [1, 2, 3, 4])), {"kind": "completeness", "columns": ["id"], "threshold": 1.0},
     status="success", metric=1.0)
case("completeness_threshold", "constraints/completeness.rs:354-374",
     tbl(email=col("i64", [1, 2, None, 4, 5])), {"kind": "completeness", "columns": ["email"], "threshold": 0.8},
     status="success", metric=0.8)
case("completeness_below_threshold", "constraints/completeness.rs:376-396",
     tbl(phone=col("i64", [1, None, None, 4])), {"kind": "completeness", "columns": ["phone"], "threshold": 0.8},
     status="failure", metric=0.5, message_contains=["50.00%"])
case("completeness_all_operator", "constraints/completeness.rs:398-421",
     tbl(first_name=col("i64", [1, 2, 3]), last_name=col("i64", [10, 20, 30])),
     {"kind": "completeness", "columns": ["first_name", "last_name"], "threshold": 1.0, "operator": ["All"]},
     status="success", metric=1.0)
case("completeness_all_operator_failure", "constraints/completeness.rs:423-448",
     tbl(col1=col("i64", [1, 2, 3]), col2=col("i64", [None, 20, 30]), col3=col("i64", [100, 200, 300])),
     {"kind": "completeness", "columns": ["col1", "col2", "col3"], "threshold": 1.0, "operator": ["All"]},
     status="failure", message_contains=["col2"])
case("completeness_any_operator", "constraints/completeness.rs:450-472",
     tbl(phone=col("i64", [1, None, None]), email=col("i64", [None, 2, None]), address=col("i64", [None, None, None])),
     {"kind": "completeness", "columns": ["phone", "email", "address"], "threshold": 0.3, "operator": ["Any"]},
     status="success")
case("completeness_at_least_operator", "constraints/completeness.rs:474-497",
     tbl(col1=col("i64", [1, 2, 3, 4]), col2=col("i64", [10, 20, 30, 40]), col3=col("i64", [None, 200, 300, 400]),
         col4=col("i64", [100, None, 3000, 4000])),
     {"kind": "completeness", "columns": ["col1", "col2", "col3", "col4"], "threshold": 0.8, "operator": ["AtLeast", 2]},
     status="success")
case("completeness_exactly_operator", "constraints/completeness.rs:499-521",
     tbl(a=col("i64", [1, 2, 3]), b=col("i64", [10, None, 30]), c=col("i64", [None, None, None])),
     {"kind": "completeness", "columns": ["a", "b", "c"], "threshold": 1.0, "operator": ["Exactly", 1]},
     status="success")
case("completeness_empty", "constraints/completeness.rs:523-533",
     tbl(id=col("i64", [])), {"kind": "completeness", "columns": ["id"], "threshold": 1.0},
     status="skipped")

# ------------------------------------------------------------------ statistics ----
V = "value"
case("stat_mean", "constraints/statistics.rs:563-573", tbl(value=col("f64", [10.0, 20.0, 30.0])),
     {"kind": "statistic", "column": V, "stat": "Mean", "assertion": ["Equals", 20.0]}, status="success", metric=20.0)
case("stat_min", "constraints/statistics.rs:575-584", tbl(value=col("f64", [5.0, 10.0, 15.0])),
     {"kind": "statistic", "column": V, "stat": "Min", "assertion": ["Equals", 5.0]}, status="success", metric=5.0)
case("stat_max", "constraints/statistics.rs:586-592", tbl(value=col("f64", [5.0, 10.0, 15.0])),
     {"kind": "statistic", "column": V, "stat": "Max", "assertion": ["Equals", 15.0]}, status="success", metric=15.0)
case("stat_sum", "constraints/statistics.rs:594-604", tbl(value=col("f64", [10.0, 20.0, 30.0])),
     {"kind": "statistic", "column": V, "stat": "Sum", "assertion": ["Equals", 60.0]}, status="success", metric=60.0)
case("stat_mean_with_nulls", "constraints/statistics.rs:606-616", tbl(value=col("f64", [10.0, None, 20.0])),
     {"kind": "statistic", "column": V, "stat": "Mean", "assertion": ["Equals", 15.0]}, status="success", metric=15.0)
case("stat_all_nulls", "constraints/statistics.rs:618-628", tbl(value=col("f64", [None, None, None])),
     {"kind": "statistic", "column": V, "stat": "Mean", "assertion": ["Equals", 0.0]},
     status="failure", message_contains=["null"])
case("multistat_success", "constraints/statistics.rs:641-660", tbl(value=col("f64", [10.0, 20.0, 30.0, 40.0])),
     {"kind": "multi_statistic", "column": V,
      "stats": [["Min", ["GreaterThanOrEqual", 10.0]], ["Max", ["LessThanOrEqual", 40.0]],
                ["Mean", ["Equals", 25.0]], ["Sum", ["Equals", 100.0]]]},
     status="success", metric=10.0)
case("multistat_failure", "constraints/statistics.rs:662-680", tbl(value=col("f64", [10.0, 20.0, 30.0])),
     {"kind": "multi_statistic", "column": V, "stats": [["Min", ["Equals", 5.0]], ["Max", ["Equals", 30.0]]]},
     status="failure", message_contains=["minimum is 10"])
# property test oracle: STDDEV is the SAMPLE statistic (tests/property_tests.rs:776-825)
case("stat_sample_stddev", "term-guard/tests/property_tests.rs:784-793", tbl(value=col("f64", [2.0, 4.0, 4.0, 4.0, 5.0, 5.0, 7.0, 9.0])),
     {"kind": "statistic", "column": V, "stat": "StandardDeviation", "assertion": ["GreaterThan", 0.0]},
     status="success", metric=2.138089935299395, metric_tol=1e-9)

# ------------------------------------------------------------------ format ----
T = "text_col"


def fmt(id, ref, values, format, threshold, expect, **kw):
    op = {"kind": "format", "column": T, "format": format, "threshold": threshold}
    op.update(kw)
    case(id, ref, tbl(text_col=col("str", values)), op, **expect)


ok75 = dict(status="success", metric=0.75)
fmt("format_email", "constraints/format.rs:916-934", ["test@example.com", "user@domain.org", "invalid-email", "another@test.net"], "Email", 0.7, ok75)
fmt("format_url", "constraints/format.rs:936-954", ["https://example.com", "http://test.org", "not-a-url", "https://another.site.net/path"], "Url", 0.7, ok75, flag=False)
fmt("format_url_localhost", "constraints/format.rs:956-973", ["https://localhost:3000", "http://localhost", "https://example.com", "not-a-url"], "Url", 0.7, ok75, flag=True)
fmt("format_credit_card_detect", "constraints/format.rs:975-994", ["4111-1111-1111-1111", "5555 5555 5555 4444", "normal text", "4111111111111111"], "CreditCard", 0.8, dict(status="success", metric=0.75), flag=True)
fmt("format_phone_us", "constraints/format.rs:996-1013", ["(555) 123-4567", "555-123-4567", "5551234567", "invalid-phone"], "Phone", 0.7, dict(status="success", metric=0.75), arg="US", trim=True)
fmt("format_postal_us", "constraints/format.rs:1015-1033", ["12345", "12345-6789", "invalid", "98765"], "PostalCode", 0.7, ok75, arg="US", trim=True)
fmt("format_uuid", "constraints/format.rs:1035-1053", ["550e8400-e29b-41d4-a716-446655440000", "6ba7b810-9dad-11d1-80b4-00c04fd430c8", "invalid-uuid", "6ba7b811-9dad-11d1-80b4-00c04fd430c8"], "UUID", 0.7, ok75)
fmt("format_ipv4", "constraints/format.rs:1055-1073", ["192.168.1.1", "10.0.0.1", "256.256.256.256", "172.16.0.1"], "IPv4", 0.7, ok75)
fmt("format_ipv6", "constraints/format.rs:1075-1093", ["2001:0db8:85a3:0000:0000:8a2e:0370:7334", "2001:db8:85a3::8a2e:370:7334", "invalid-ipv6", "::1"], "IPv6", 0.7, ok75)
fmt("format_json", "constraints/format.rs:1095-1113", ['{"key": "value"}', "[1, 2, 3]", "not json", '{"nested": {"key": "value"}}'], "Json", 0.7, ok75)
fmt("format_iso8601", "constraints/format.rs:1115-1133", ["2023-12-25T10:30:00Z", "2023-12-25T10:30:00.123Z", "invalid-datetime", "2023-12-25T10:30:00+05:30"], "Iso8601DateTime", 0.7, ok75)
fmt("format_custom_regex", "constraints/format.rs:1135-1155", ["ABC123", "DEF456", "invalid", "GHI789"], "Regex", 0.7, ok75, arg=r"^[A-Z]{3}\d{3}$")
fmt("format_case_insensitive", "constraints/format.rs:1157-1182", ["abc123", "DEF456", "invalid", "ghi789"], "Regex", 0.7, ok75, arg=r"^[A-Z]{3}\d{3}$", case_sensitive=False)
fmt("format_trim", "constraints/format.rs:1184-1208", ["  test@example.com  ", "user@domain.org", "  invalid-email  ", " another@test.net "], "Email", 0.7, ok75, trim=True)
fmt("format_null_valid", "constraints/format.rs:1210-1228", ["test@example.com", None, "invalid-email", None], "Email", 0.6, ok75, null_is_valid=True)
fmt("format_null_invalid", "constraints/format.rs:1230-1245", ["test@example.com", None, "invalid-email", None], "Email", 0.2, dict(status="success", metric=0.25), null_is_valid=False)
fmt("format_failure", "constraints/format.rs:1247-1265", ["invalid", "also_invalid", "nope", "still_invalid"], "Email", 0.5, dict(status="failure", metric=0.0, message_contains=["Format validation ratio 0.000 is below threshold 0.500"]))
fmt("format_empty", "constraints/format.rs:1267-1277", [], "Email", 0.9, dict(status="skipped"))
fmt("ssn_valid", "constraints/format.rs:1389-1407", ["123-45-6789", "123456789", "456-78-9012", "789012345"], "SocialSecurityNumber", 0.95, dict(status="success", metric=1.0), trim=True)
fmt("ssn_invalid", "constraints/format.rs:1409-1431", ["000-12-3456", "666-12-3456", "900-12-3456", "123-00-4567", "123-45-0000"], "SocialSecurityNumber", 0.0, dict(status="success", metric=0.0), trim=True)
fmt("ssn_mixed", "constraints/format.rs:1433-1458", ["123-45-6789", "not-an-ssn", "666-12-3456", "456789012", "123 45 6789", "789-01-2345", None, "234-56-7890"], "SocialSecurityNumber", 0.5, dict(status="success", metric=0.625), trim=True)
fmt("ssn_threshold_fail", "constraints/format.rs:1460-1475", ["123-45-6789", "invalid", "234-56-7890", "not-ssn"], "SocialSecurityNumber", 0.8, dict(status="failure", metric=0.5), trim=True)
fmt("ssn_threshold_pass", "constraints/format.rs:1477-1484", ["123-45-6789", "invalid", "234-56-7890", "not-ssn"], "SocialSecurityNumber", 0.4, dict(status="success", metric=0.5), trim=True)
fmt("ssn_edge_cases", "constraints/format.rs:1486-1508", ["078-05-1120", "219-09-9999", "457-55-5462", "999-99-9999", "123-45-67890", "12-345-6789", "ABC-DE-FGHI", ""], "SocialSecurityNumber", 0.3, dict(status="success", metric=0.375), trim=True)

# ------------------------------------------------------------------ uniqueness ----
U = "test_col"


def uniq(id, ref, values, kind, expect, **kw):
    op = {"kind": "uniqueness", "columns": [U], "uniqueness": kind}
    op.update(kw)
    case(id, ref, tbl(test_col=col("str", values)), op, **expect)


uniq("uniq_full_single", "constraints/uniqueness.rs:907-920", ["A", "B", "C", "A"], "FullUniqueness", dict(status="success", metric=0.75), threshold=0.7)
uniq("uniq_full_with_nulls", "constraints/uniqueness.rs:922-936", ["A", "B", None, "A"], "FullUniqueness", dict(status="success", metric=0.5), threshold=0.4)
uniq("uniq_distinctness", "constraints/uniqueness.rs:938-951", ["A", "B", "C", "A"], "Distinctness", dict(status="success", metric=0.75), assertion=["Equals", 0.75])
uniq("uniq_unique_value_ratio", "constraints/uniqueness.rs:953-967", ["A", "B", "C", "A"], "UniqueValueRatio", dict(status="success", metric=0.5), assertion=["Equals", 0.5])
uniq("uniq_pk_success", "constraints/uniqueness.rs:969-980", ["A", "B", "C"], "PrimaryKey", dict(status="success", metric=1.0))
uniq("uniq_pk_nulls", "constraints/uniqueness.rs:982-994", ["A", "B", None], "PrimaryKey", dict(status="failure", message_contains=["NULL values"]))
uniq("uniq_pk_duplicates", "constraints/uniqueness.rs:996-1006", ["A", "B", "A"], "PrimaryKey", dict(status="failure", message_contains=["duplicate values"]))
uniq("uniq_with_nulls_include", "constraints/uniqueness.rs:1042-1057", ["A", "B", None, None], "UniqueWithNulls", dict(status="success", metric=0.75), threshold=0.4, null_handling="Include")
uniq("uniq_empty", "constraints/uniqueness.rs:1059-1070", [], "FullUniqueness", dict(status="skipped"), threshold=1.0)
case("uniq_multi_full", "constraints/uniqueness.rs:1008-1022",
     tbl(col1=col("str", ["A", "B", "A"]), col2=col("str", ["1", "2", "2"])),
     {"kind": "uniqueness", "columns": ["col1", "col2"], "uniqueness": "FullUniqueness", "threshold": 0.9},
     status="success", metric=1.0)
case("uniq_multi_distinctness", "constraints/uniqueness.rs:1024-1040",
     tbl(col1=col("str", ["A", "B", "A"]), col2=col("str", ["1", "2", "1"])),
     {"kind": "uniqueness", "columns": ["col1", "col2"], "uniqueness": "Distinctness", "assertion": ["GreaterThan", 0.5]},
     status="success", metric=2.0 / 3.0, metric_tol=0.01)

# ------------------------------------------------------------------ correlation ----
xs = [float(i) for i in range(100)]
corr_tbl = tbl(x=col("f64", xs), y=col("f64", [2.0 * i + (i % 10) - 5.0 for i in range(100)]))
ind_tbl = tbl(x=col("f64", xs), y=col("f64", [float((i * 37) % 100) for i in range(100)]))
case("corr_pearson", "constraints/correlation.rs:587-598", corr_tbl,
     {"kind": "correlation", "c1": "x", "c2": "y", "corr": "Pearson", "assertion": ["GreaterThan", 0.9]},
     status="success", metric_gt=0.9)
case("corr_independence", "constraints/correlation.rs:600-612", ind_tbl,
     {"kind": "correlation", "c1": "x", "c2": "y", "corr": "Independence", "assertion": ["Equals", 0.3]},
     status="success")
case("corr_range", "constraints/correlation.rs:614-631", corr_tbl,
     {"kind": "correlation", "c1": "x", "c2": "y", "corr": "Range", "assertion": ["Between", 0.8, 1.0]},
     status="success")
case("corr_spearman_skipped", "constraints/correlation.rs:340-345", corr_tbl,
     {"kind": "correlation", "c1": "x", "c2": "y", "corr": "Spearman", "assertion": ["GreaterThan", 0.9]},
     status="skipped", message_contains=["Correlation type not yet implemented"])

# ------------------------------------------------------------------ custom sql ----
sql_tbl = tbl(price=col("f64", [10.5, 25.0, 5.0, 100.0, None]), quantity=col("i64", [5, 10, 0, 20, 15]),
              status=col("str", ["active", "active", "inactive", "active", "pending"]))


def sql(id, ref, expr, expect, hint=None):
    case(id, ref, sql_tbl, {"kind": "custom_sql", "expression": expr, "hint": hint}, **expect)


sql("sql_nulls_expression", "constraints/custom_sql.rs:394-405", "price > 0", dict(status="failure", metric=0.8))
sql("sql_all_satisfy", "constraints/custom_sql.rs:407-419", "quantity >= 0", dict(status="success", metric=1.0))
sql("sql_partial_satisfy", "constraints/custom_sql.rs:421-438", "quantity > 0", dict(status="failure", metric=0.8, message_contains=["Quantity must be positive", "1 rows failed"]), hint="Quantity must be positive")
sql("sql_complex", "constraints/custom_sql.rs:440-456", "status = 'active' AND price >= 10", dict(status="failure", metric=0.6), hint="Active items must have price >= 10")
sql("sql_is_not_null", "constraints/custom_sql.rs:458-469", "price IS NOT NULL", dict(status="failure", metric=0.8))
sql("sql_invalid_column", "constraints/custom_sql.rs:471-486", "invalid_column > 0", dict(status="failure", message_contains=["SQL expression error", "invalid_column"]))

# ------------------------------------------------------------------ length ----
def scol(v):
    return {"data": {"text": col("str", v)}}


case("length_min_ok", "constraints/length.rs:244-267", scol(["hello", "world", "testing", "great", None]),
     {"kind": "length", "column": "text", "assertion": ["Min", 5]}, status="success", metric=1.0)
case("length_min_failure", "constraints/length.rs:269-288", scol(["hi", "hello", "a", "testing", None]),
     {"kind": "length", "column": "text", "assertion": ["Min", 5]}, status="failure", metric=0.6,
     message_contains=["at least 5 characters"])
case("length_max_ok", "constraints/length.rs:290-302", scol(["hi", "hey", "test", None]),
     {"kind": "length", "column": "text", "assertion": ["Max", 10]}, status="success", metric=1.0)
case("length_max_failure", "constraints/length.rs:304-322",
     scol(["short", "this is a very long string that exceeds the limit", "ok", None]),
     {"kind": "length", "column": "text", "assertion": ["Max", 10]}, status="failure", metric=0.75,
     message_contains=["at most 10 characters"])
case("length_between", "constraints/length.rs:324-347", scol(["hello", "testing", "hi", "this is way too long", None]),
     {"kind": "length", "column": "text", "assertion": ["Between", 3, 10]}, status="failure", metric=0.6,
     message_contains=["between 3 and 10 characters"])
case("length_exactly", "constraints/length.rs:349-369", scol(["hello", "world", "test", "testing", None]),
     {"kind": "length", "column": "text", "assertion": ["Exactly", 5]}, status="failure", metric=0.6,
     message_contains=["exactly 5 characters"])
case("length_not_empty", "constraints/length.rs:371-391", scol(["hello", "a", "", "testing", None]),
     {"kind": "length", "column": "text", "assertion": ["NotEmpty"]}, status="failure", metric=0.8,
     message_contains=["not empty"])
case("length_utf8_characters_not_bytes", "constraints/length.rs:393-412", scol(["hello", "你好", "🦀🔥", "café", None]),
     {"kind": "length", "column": "text", "assertion": ["Min", 2]}, status="success")
case("length_all_null", "constraints/length.rs:414-426", scol([None, None, None]),
     {"kind": "length", "column": "text", "assertion": ["Min", 5]}, status="success", metric=1.0)
case("length_empty", "constraints/length.rs:428-438", scol([]),
     {"kind": "length", "column": "text", "assertion": ["Min", 5]}, status="skipped")

# ------------------------------------------------------------------ containment / non-negative ----
case("containment_failure", "constraints/values.rs:523-543",
     {"data": {"text_col": col("str", ["active", "inactive", "pending", "invalid_status"])}},
     {"kind": "containment", "column": "text_col", "allowed": ["active", "inactive", "pending", "archived"]},
     status="failure", metric=0.75)
case("containment_ok", "constraints/values.rs:545-560",
     {"data": {"text_col": col("str", ["active", "inactive", "pending"])}},
     {"kind": "containment", "column": "text_col", "allowed": ["active", "inactive", "pending", "archived"]},
     status="success", metric=1.0)
case("containment_with_nulls", "constraints/values.rs:590-602",
     {"data": {"text_col": col("str", ["active", None, "inactive", None])}},
     {"kind": "containment", "column": "text_col", "allowed": ["active", "inactive"]}, status="success", metric=1.0)
case("non_negative_ok", "constraints/values.rs:562-574", {"data": {"num_col": col("f64", [1.0, 0.0, 5.5, 100.0])}},
     {"kind": "non_negative", "column": "num_col"}, status="success", metric=1.0)
case("non_negative_failure", "constraints/values.rs:576-588", {"data": {"num_col": col("f64", [1.0, -2.0, 5.5, 100.0])}},
     {"kind": "non_negative", "column": "num_col"}, status="failure", metric=0.75)

# ------------------------------------------------------------------ data type ----
case("data_type_integer", "constraints/values.rs:480-493", {"data": {"text_col": col("str", ["123", "456", "not_number", "789"])}},
     {"kind": "data_type", "column": "text_col", "data_type": "Integer", "threshold": 0.7}, status="success", metric=0.75)
case("data_type_float", "constraints/values.rs:495-507", {"data": {"text_col": col("str", ["123.45", "67.89", "invalid", "100"])}},
     {"kind": "data_type", "column": "text_col", "data_type": "Float", "threshold": 0.7}, status="success", metric=0.75)
case("data_type_boolean", "constraints/values.rs:509-521", {"data": {"text_col": col("str", ["true", "false", "invalid", "1"])}},
     {"kind": "data_type", "column": "text_col", "data_type": "Boolean", "threshold": 0.7}, status="success", metric=0.75)

# ------------------------------------------------------------------ column count ----
def ncols(k):
    return {"data": {f"col_{i}": col("i64", [i, None]) for i in range(k)}}


case("column_count_equals", "constraints/column_count.rs:138-149", ncols(5), {"kind": "column_count", "assertion": ["Equals", 5.0]},
     status="success", metric=5.0)
case("column_count_equals_failure", "constraints/column_count.rs:150-158", ncols(5), {"kind": "column_count", "assertion": ["Equals", 10.0]},
     status="failure", metric=5.0, message_contains=["Column count 5 does not satisfy assertion equals 10"])
case("column_count_greater_than", "constraints/column_count.rs:160-178", ncols(8), {"kind": "column_count", "assertion": ["GreaterThan", 10.0]},
     status="failure", metric=8.0)
case("column_count_between", "constraints/column_count.rs:198-215", ncols(7), {"kind": "column_count", "assertion": ["Between", 5.0, 10.0]},
     status="success", metric=7.0)
case("column_count_large", "constraints/column_count.rs:229-240", ncols(100), {"kind": "column_count", "assertion": ["GreaterThanOrEqual", 100.0]},
     status="success", metric=100.0)

# ------------------------------------------------------------------ approx_count_distinct ----
case("approx_distinct_high_cardinality", "constraints/approx_count_distinct.rs:190-205",
     {"data": {"test_col": col("i64", list(range(1000)))}},
     {"kind": "approx_count_distinct", "column": "test_col", "assertion": ["GreaterThan", 990.0]},
     status="success", metric_gt=990.0)
case("approx_distinct_low_cardinality", "constraints/approx_count_distinct.rs:207-226",
     {"data": {"test_col": col("i64", [1, 2, 3] * 100)}},
     {"kind": "approx_count_distinct", "column": "test_col", "assertion": ["LessThan", 10.0]},
     status="success", metric_lt=10.0)
case("approx_distinct_with_nulls", "constraints/approx_count_distinct.rs:228-255",
     {"data": {"test_col": col("i64", [1, None, 2, None, 3, None, 1, 2, 3, None])}},
     {"kind": "approx_count_distinct", "column": "test_col", "assertion": ["Between", 2.0, 5.0]},
     status="success", metric_gt=1.999, metric_lt=5.001)
case("approx_distinct_failure", "constraints/approx_count_distinct.rs:257-273",
     {"data": {"test_col": col("i64", [i % 10 for i in range(50)])}},
     {"kind": "approx_count_distinct", "column": "test_col", "assertion": ["GreaterThan", 100.0]},
     status="failure", metric_lt=20.0)
case("approx_distinct_strings", "constraints/approx_count_distinct.rs:275-297",
     {"data": {"test_col": col("str", ["apple", "banana", "cherry", "apple", "banana", "date", "elderberry", None])}},
     {"kind": "approx_count_distinct", "column": "test_col", "assertion": ["Between", 4.0, 6.0]}, status="success")
case("approx_distinct_empty", "constraints/approx_count_distinct.rs:299-311",
     {"data": {"test_col": col("i64", [])}},
     {"kind": "approx_count_distinct", "column": "test_col", "assertion": ["Equals", 0.0]}, status="success", metric=0.0)
case("approx_distinct_all_null", "constraints/approx_count_distinct.rs:313-326",
     {"data": {"test_col": col("i64", [None, None, None, None, None])}},
     {"kind": "approx_count_distinct", "column": "test_col", "assertion": ["Equals", 0.0]}, status="success", metric=0.0)

# ------------------------------------------------------------------ quantile constraint ----
_q100 = {"data": {"value": col("f64", [float(i) for i in range(1, 101)])}}
case("quantile_median", "constraints/quantile.rs:526-538", _q100,
     {"kind": "quantile", "column": "value", "mode": "Single", "checks": [[0.5, ["Between", 45.0, 55.0]]]},
     status="success", metric_gt=44.999, metric_lt=55.001)
case("quantile_percentile_95", "constraints/quantile.rs:540-552", _q100,
     {"kind": "quantile", "column": "value", "mode": "Single", "checks": [[0.95, ["Between", 94.0, 96.0]]]},
     status="success", metric_gt=93.999, metric_lt=96.001)
case("quantile_multiple", "constraints/quantile.rs:554-572", _q100,
     {"kind": "quantile", "column": "value", "mode": "Multiple",
      "checks": [[0.25, ["Between", 24.0, 26.0]], [0.75, ["Between", 74.0, 76.0]]]}, status="success")
case("quantile_monotonic_strict", "constraints/quantile.rs:574-592", _q100,
     {"kind": "quantile", "column": "value", "mode": "Monotonic", "quantiles": [0.1, 0.5, 0.9], "strict": True}, status="success")

# ------------------------------------------------------------------ foreign key ----
def fk(id, ref, parent_ids, child_ids, expect, allow_nulls=False):
    case(id, ref, {"customers": {"id": col("i64", parent_ids)}, "orders": {"customer_id": col("i64", child_ids)}},
         {"kind": "foreign_key", "child": "orders.customer_id", "parent": "customers.id", "allow_nulls": allow_nulls}, **expect)


fk("fk_success", "constraints/foreign_key.rs:423-452", [1, 2], [1, 2], dict(status="success", message_none=True))
fk("fk_violation", "constraints/foreign_key.rs:454-492", [1, 2, 3], [1, 2, 999, 998],
   dict(status="failure", metric=2.0, message_contains=["Foreign key constraint violation", "2 values", "orders.customer_id", "customers.id"]))
fk("fk_nulls_disallowed", "constraints/foreign_key.rs:494-527", [1], [1, None], dict(status="failure"))
fk("fk_nulls_allowed", "constraints/foreign_key.rs:529-561", [1], [1, None], dict(status="success"), allow_nulls=True)
# fixture orders_with_orphans (test_fixtures.rs:374-438): product_id 6 and 7 orphaned
fk("fk_orphans_fixture", "test_fixtures.rs:374-438", [1, 2, 3, 4, 5], [1, 2, 3, 6, 4, 5, 7, 1], dict(status="failure", metric=2.0))

# ------------------------------------------------------------------ size ----
case("size_equals", "term-guard/tests/property_tests.rs:309-365", tbl(id=col("i64", [1, 2, 3, 4, None])),
     {"kind": "size", "assertion": ["Equals", 5.0]}, status="success", metric=5.0)
case("size_empty_equals_zero", "term-guard/tests/integration_test_suite.rs:441-466", tbl(id=col("i64", [])),
     {"kind": "size", "assertion": ["Equals", 0.0]}, status="success", metric=0.0)
case("size_failure_message", "constraints/size.rs:101-116", tbl(id=col("i64", [1, 2, 3])),
     {"kind": "size", "assertion": ["GreaterThan", 10.0]}, status="failure", metric=3.0,
     message_contains=["Size 3 does not greater than 10"])

# the constraint's own tests: value = 0 .. n-1 (constraints/size.rs:147-167)
for nm, src, n, assertion, status in (("size_rs_equals", "constraints/size.rs:169-179", 100, ["Equals", 100.0], "success"),
                                      ("size_rs_greater_than", "constraints/size.rs:181-191", 50, ["GreaterThan", 25.0], "success"),
                                      ("size_rs_between", "constraints/size.rs:193-203", 75, ["Between", 50.0, 100.0], "success"),
                                      ("size_rs_failure", "constraints/size.rs:205-215", 10, ["GreaterThan", 50.0], "failure"),
                                      ("size_rs_empty", "constraints/size.rs:217-227", 0, ["Equals", 0.0], "success")):
    case(nm, src, tbl(value=col("i64", list(range(n)))), {"kind": "size", "assertion": assertion}, status=status, metric=float(n))

# ------------------------------------------------------------------ analyzers ----
an_tbl = tbl(id=col("i64", [1, 2, 3, 4, None]), value=col("f64", [10.0, 20.0, None, 30.0, 40.0]),
             name=col("str", ["a", "b", "a", None, "c"]))


def an(id, ref, op, tables=None, **expect):
    case(id, ref, tables or an_tbl, dict(kind="analyzer", **op), **expect)


an("an_size", "analyzers/basic/tests.rs:43-52", {"analyzer": "Size"}, u=[5], metric_long=5)
an("an_completeness_id", "analyzers/basic/tests.rs:72-91", {"analyzer": "Completeness", "column": "id"}, u=[5, 4], metric=0.8)
an("an_distinctness_name", "analyzers/basic/tests.rs:116-126", {"analyzer": "Distinctness", "column": "name"}, u=[4, 3], metric=0.75)
an("an_mean", "analyzers/basic/tests.rs:135-146", {"analyzer": "Mean", "column": "value"}, u=[4], f=[100.0], metric=25.0)
an("an_min", "analyzers/basic/tests.rs:170-180", {"analyzer": "Min", "column": "value"}, f=[10.0, 40.0], metric=10.0)
an("an_max", "analyzers/basic/tests.rs:182-192", {"analyzer": "Max", "column": "value"}, f=[10.0, 40.0], metric=40.0)
an("an_sum", "analyzers/basic/tests.rs:217-228", {"analyzer": "Sum", "column": "value"}, f=[100.0], metric=100.0)
empty_tbl = tbl(value=col("f64", []))
an("an_empty_size", "analyzers/basic/tests.rs:263-273", {"analyzer": "Size"}, tables=empty_tbl, u=[0], metric_long=0)
an("an_empty_completeness", "analyzers/basic/tests.rs:275-279", {"analyzer": "Completeness", "column": "value"}, tables=empty_tbl, metric=1.0)
an("an_empty_mean", "analyzers/basic/tests.rs:281-285", {"analyzer": "Mean", "column": "value"}, tables=empty_tbl, no_data=True)
null_tbl = tbl(value=col("f64", [None, None, None]))
an("an_nullonly_size", "analyzers/basic/tests.rs:300-306", {"analyzer": "Size"}, tables=null_tbl, u=[3], metric_long=3)
an("an_nullonly_completeness", "analyzers/basic/tests.rs:308-312", {"analyzer": "Completeness", "column": "value"}, tables=null_tbl, metric=0.0)
an("an_nullonly_sum", "analyzers/basic/tests.rs:314-320", {"analyzer": "Sum", "column": "value"}, tables=null_tbl, no_data=True)
lin_tbl = tbl(x=col("f64", xs), y=col("f64", [2.0 * v + 1.0 for v in xs]))
an("an_pearson_perfect", "analyzers/advanced/correlation.rs:497-510", {"analyzer": "Pearson", "column": "x", "column2": "y"}, tables=lin_tbl, metric=1.0, metric_tol=1e-4)
an("an_covariance", "analyzers/advanced/correlation.rs:512-530", {"analyzer": "Covariance", "column": "x", "column2": "y"}, tables=lin_tbl, metric_gt=1600.0, metric_lt=1700.0)
an("an_spearman", "analyzers/advanced/correlation.rs:532-548", {"analyzer": "Spearman", "column": "x", "column2": "y"}, tables=lin_tbl, metric=1.0, metric_tol=1e-4)
# grouped completeness (analyzers/basic/grouped_completeness.rs:309-365): 4 groups; US|A 1.0, EU|A 0.5
grp_tbl = tbl(region=col("str", ["US", "US", "EU", "EU", "US", "EU"]), product=col("str", ["A", "B", "A", "B", "A", "A"]),
              sales=col("i32", [100, 200, None, 150, 250, 300]))
an("an_grouped_completeness", "analyzers/basic/grouped_completeness.rs:309-365",
   {"analyzer": "GroupedCompleteness", "column": "sales", "groups": ["region", "product"]}, tables=grp_tbl,
   map={"US_A": 1.0, "EU_A": 0.5, "US_B": 1.0, "EU_B": 1.0}, n_groups=4)

# ------------------------------------------------------------------ assertion / logical ----
case("assertion_cases", "constraints/assertion.rs:88-128", {}, {"kind": "assertion", "cases": [
    [["Equals", 10.0], 10.0, True], [["Equals", 10.0], 10.1, False], [["NotEquals", 10.0], 10.0, False],
    [["NotEquals", 10.0], 10.1, True], [["GreaterThan", 10.0], 10.1, True], [["GreaterThan", 10.0], 10.0, False],
    [["GreaterThan", 10.0], 9.9, False], [["Between", 10.0, 20.0], 15.0, True], [["Between", 10.0, 20.0], 10.0, True],
    [["Between", 10.0, 20.0], 20.0, True], [["Between", 10.0, 20.0], 9.9, False], [["Between", 10.0, 20.0], 20.1, False]],
    "descriptions": [[["Equals", 10.0], "equals 10"], [["GreaterThan", 5.0], "greater than 5"], [["Between", 1.0, 10.0], "between 1 and 10"]]})
case("logical_cases", "core/logical.rs:308-369", {}, {"kind": "logical", "cases": [
    [["All"], [True, True, True], True], [["All"], [True, False, True], False], [["All"], [], True],
    [["Any"], [False, True, False], True], [["Any"], [False, False], False], [["Any"], [], False],
    [["Exactly", 2], [True, True, False], True], [["Exactly", 2], [True, False, False], False], [["Exactly", 0], [], True],
    [["AtLeast", 2], [True, True, True], True], [["AtLeast", 2], [True, False, False], False],
    [["AtMost", 1], [True, False, False], True], [["AtMost", 1], [True, True, False], False], [["AtMost", 0], [], True]]})

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(out, "w") as f:
        json.dump(C, f, indent=1)
    print(f"wrote {len(C)} cases to {out}")
