"""Shared test plumbing: golden-case loading, table construction, and the two evaluators
(oracle = checker, term_b200 = product through the C ABI)."""
import json
import os

import numpy as np
import pyarrow as pa

HERE = os.path.dirname(os.path.abspath(__file__))
PA_TYPES = {"i64": pa.int64(), "f64": pa.float64(), "str": pa.string(), "i32": pa.int32(), "f32": pa.float32(), "bool": pa.bool_()}


def load_golden():
    with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
        return json.load(f)


def arrow_table(cols: dict) -> pa.Table:
    arrays, names = [], []
    for name, c in cols.items():
        arrays.append(pa.array(c["values"], type=PA_TYPES[c["type"]]))
        names.append(name)
    return pa.Table.from_arrays(arrays, names=names)


def case_tables(case):
    return {name: arrow_table(cols) for name, cols in case["tables"].items()}


# ---------------------------------------------------------------- oracle side ----
def oracle_eval(case):
    from oracle import term_oracle as O
    op = case["op"]
    tables = case_tables(case)
    t = tables.get("data")
    k = op["kind"]
    if k == "completeness":
        return O.completeness(t, op["columns"], op["threshold"], tuple(op.get("operator", ["All"])))
    if k == "size":
        return O.size(t, tuple(op["assertion"]))
    if k == "statistic":
        return O.statistic(t, op["column"], op["stat"], tuple(op["assertion"]))
    if k == "multi_statistic":
        return O.multi_statistic(t, op["column"], [(s, tuple(a)) for s, a in op["stats"]])
    if k == "format":
        return O.format_constraint(t, op["column"], op["format"], op["threshold"], op.get("arg"), op.get("flag", False),
                                   op.get("case_sensitive", True), op.get("trim", False), op.get("null_is_valid", True))
    if k == "uniqueness":
        return O.uniqueness(t, op["columns"], op["uniqueness"], op.get("threshold", 1.0),
                            tuple(op["assertion"]) if "assertion" in op else None, op.get("null_handling", "Exclude"))
    if k == "correlation":
        return O.correlation(t, op["c1"], op["c2"], "Pearson" if op["corr"] == "Range" else op["corr"], tuple(op["assertion"]))
    if k == "custom_sql":
        try:
            return O.custom_sql(t, op["expression"], op.get("hint"))
        except KeyError as e:
            return O.Result("failure", None, f"SQL expression error: Schema error: No field named {e.args[0]}. Expression: '{op['expression']}'")
    if k == "foreign_key":
        return O.foreign_key(tables, op["child"], op["parent"], op.get("allow_nulls", False))[0]
    if k == "length":
        return O.length_constraint(t, op["column"], op["assertion"][0], *op["assertion"][1:])
    if k == "containment":
        return O.containment(t, op["column"], op["allowed"])
    if k == "non_negative":
        return O.non_negative(t, op["column"])
    if k == "approx_count_distinct":
        return O.approx_count_distinct(t, op["column"], tuple(op["assertion"]))
    if k == "quantile":
        return O.quantile_constraint(t, op["column"], op["mode"], checks=[(q, tuple(a)) for q, a in op.get("checks", [])],
                                     quantiles=op.get("quantiles", ()), strict=op.get("strict", False))
    if k == "data_type":
        return O.data_type(t, op["column"], op["data_type"], op["threshold"])
    if k == "column_count":
        return O.column_count(t, tuple(op["assertion"]))
    raise ValueError(k)


def oracle_analyzer(case):
    """returns dict(u=[..], f=[..], metric=..., metric_long=..., no_data=bool, map={})"""
    from oracle import term_oracle as O
    op = case["op"]
    t = case_tables(case)["data"]
    a = op["analyzer"]
    cols = O.table_cols(t)
    if a == "Size":
        n = O.n_rows(cols)
        return dict(u=[n], metric_long=n)
    if a == "Completeness":
        tt, nn, m = O.an_completeness(t, op["column"])
        return dict(u=[tt, nn], metric=m)
    if a == "Distinctness":
        nn, d, m = O.an_distinctness(t, op["column"])
        return dict(u=[nn, d], metric=m)
    c = cols.get(op.get("column"))
    if a in ("Mean", "Sum", "Min", "Max"):
        v = c.values[c.valid]
        if len(v) == 0:
            return dict(no_data=True)
        import math
        s = math.fsum(float(x) for x in v)
        if a == "Mean":
            return dict(u=[len(v)], f=[s], metric=s / len(v))
        if a == "Sum":
            return dict(f=[s], metric=s)
        return dict(f=[float(v.min()), float(v.max())], metric=float(v.min() if a == "Min" else v.max()))
    if a in ("Pearson", "Covariance", "Spearman"):
        return dict(metric=O.an_correlation(t, op["column"], op["column2"], a.lower()))
    if a == "GroupedCompleteness":
        g = O.grouped_completeness(t, op["column"], op["groups"])
        return dict(map={"_".join(k): (nn / tt if tt else 1.0) for k, (tt, nn) in g.items()}, n_groups=len(g))
    raise ValueError(a)


# ---------------------------------------------------------------- product side (C ABI) ----
def _assertion(T, a):
    return getattr(T.Assertion, a[0])(*a[1:])


def _operator(T, o):
    if o[0] == "All":
        return T.LogicalOperator.All
    if o[0] == "Any":
        return T.LogicalOperator.Any
    return getattr(T.LogicalOperator, o[0])(o[1])


def build_constraint(T, op):
    k = op["kind"]
    if k == "completeness":
        return T.CompletenessConstraint(op["columns"], op["threshold"], _operator(T, op.get("operator", ["All"])))
    if k == "size":
        return T.SizeConstraint(_assertion(T, op["assertion"]))
    if k == "statistic":
        return T.StatisticalConstraint(op["column"], T.StatisticType[op["stat"]], _assertion(T, op["assertion"]))
    if k == "multi_statistic":
        return T.MultiStatisticalConstraint(op["column"], [(T.StatisticType[s], _assertion(T, a)) for s, a in op["stats"]])
    if k == "format":
        opts = T.FormatOptions(op.get("case_sensitive", True), op.get("trim", False), op.get("null_is_valid", True))
        return T.FormatConstraint(op["column"], T.FormatType[op["format"]], op["threshold"], opts, op.get("arg"), op.get("flag", False))
    if k == "uniqueness":
        return T.UniquenessConstraint(op["columns"], T.UniquenessType[op["uniqueness"]], op.get("threshold", 1.0),
                                      _assertion(T, op["assertion"]) if "assertion" in op else None,
                                      T.NullHandling[op.get("null_handling", "Exclude")])
    if k == "correlation":
        return T.CorrelationConstraint(op["c1"], op["c2"], T.CorrelationType[op["corr"]], _assertion(T, op["assertion"]))
    if k == "custom_sql":
        return T.CustomSqlConstraint(op["expression"], op.get("hint"))
    if k == "foreign_key":
        return T.ForeignKeyConstraint(op["child"], op["parent"]).allow_nulls(op.get("allow_nulls", False))
    if k == "length":
        return T.LengthConstraint(op["column"], getattr(T.LengthAssertion, op["assertion"][0])(*op["assertion"][1:]))
    if k == "containment":
        return T.ContainmentConstraint(op["column"], op["allowed"])
    if k == "non_negative":
        return T.NonNegativeConstraint(op["column"])
    if k == "approx_count_distinct":
        return T.ApproxCountDistinctConstraint(op["column"], _assertion(T, op["assertion"]))
    if k == "quantile":
        if op["mode"] == "Monotonic":
            return T.QuantileConstraint.monotonic(op["column"], op["quantiles"], op["strict"])
        checks = [T.QuantileCheck(q, _assertion(T, a)) for q, a in op["checks"]]
        return T.QuantileConstraint(op["column"], T.QuantileConstraint.SINGLE if op["mode"] == "Single" else T.QuantileConstraint.MULTIPLE, checks)
    if k == "data_type":
        return T.DataTypeConstraint(op["column"], T.DataType[op["data_type"]], op["threshold"])
    if k == "column_count":
        return T.ColumnCountConstraint(_assertion(T, op["assertion"]))
    raise ValueError(k)


def build_analyzer(T, op):
    a = op["analyzer"]
    if a == "Size":
        return T.SizeAnalyzer()
    simple = {"Completeness": T.CompletenessAnalyzer, "Distinctness": T.DistinctnessAnalyzer, "Mean": T.MeanAnalyzer,
              "Min": T.MinAnalyzer, "Max": T.MaxAnalyzer, "Sum": T.SumAnalyzer, "StandardDeviation": T.StandardDeviationAnalyzer}
    if a in simple:
        return simple[a](op["column"])
    if a == "Pearson":
        return T.CorrelationAnalyzer.pearson(op["column"], op["column2"])
    if a == "Covariance":
        return T.CorrelationAnalyzer.covariance(op["column"], op["column2"])
    if a == "Spearman":
        return T.CorrelationAnalyzer.spearman(op["column"], op["column2"])
    if a == "GroupedCompleteness":
        return T.GroupedCompletenessAnalyzer(op["column"], op["groups"])
    raise ValueError(a)


_counter = [0]


def register_case(ctx, case, prefix=None):
    """Registers the case's tables under unique names; returns mapping original -> registered name."""
    names = {}
    for name, cols in case["tables"].items():
        _counter[0] += 1
        reg = name if prefix is None else f"{prefix}{_counter[0]}_{name}"
        ctx.register_table(reg, arrow_table(cols))
        names[name] = reg
    return names


def check_expect(result_status, metric, message, expect, what=""):
    assert result_status == expect["status"], f"{what}: status {result_status} != {expect['status']} ({message})"
    if "metric" in expect:
        assert metric is not None, f"{what}: metric is None"
        tol = expect.get("metric_tol", 0.0)
        if tol:
            assert abs(metric - expect["metric"]) <= tol, f"{what}: metric {metric} vs {expect['metric']}"
        else:
            assert metric == expect["metric"], f"{what}: metric {metric!r} != {expect['metric']!r}"
    if "metric_gt" in expect:
        assert metric is not None and metric > expect["metric_gt"], f"{what}: metric {metric}"
    if "metric_lt" in expect:
        assert metric is not None and metric < expect["metric_lt"], f"{what}: metric {metric}"
    for frag in expect.get("message_contains", []):
        assert message is not None and frag in message, f"{what}: message {message!r} lacks {frag!r}"
    if expect.get("message_none"):
        assert message is None, f"{what}: message {message!r}"


# ---- numpy restatement of the key hash of term_b200/csrc/hash_common.cuh (test infrastructure) ----
def fmix64_np(k):
    import numpy as np
    k = np.asarray(k).astype(np.uint64)
    with np.errstate(over="ignore"):
        k = k ^ (k >> np.uint64(33))
        k = k * np.uint64(0xff51afd7ed558ccd)
        k = k ^ (k >> np.uint64(33))
        k = k * np.uint64(0xc4ceb9fe1a85ec53)
        k = k ^ (k >> np.uint64(33))
    return k


def hash_rank_np(keys_i64, world):
    """destination rank of the multi-GPU shuffle: ((hash >> 42) * world) >> 22"""
    import numpy as np
    h = fmix64_np(np.asarray(keys_i64).view(np.uint64))
    return (((h >> np.uint64(42)) * np.uint64(world)) >> np.uint64(22)).astype(np.int64)
