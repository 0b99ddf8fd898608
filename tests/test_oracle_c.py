"""oracle/cpu_scan.c — the C restatement behind bench.py's cpu_baseline and `--impl reference` legs, i.e. the
denominator of every GPU / CPU ratio — pinned against the numpy oracle (which is pinned against the reference's golden
vectors) on seeded data, including the edge cases the reference's SQL has: all-NULL columns, no validity bitmap at all,
n = 0 / 1 / 2 rows, ragged sizes around the 64-bit popcount words."""
import math

import numpy as np
import pyarrow as pa
import pytest

from oracle import cpu_scan as S
from oracle import term_oracle as O


def _table(n, seed, null_frac=0.05, all_null=()):
    rng = np.random.default_rng(seed)
    f0 = rng.normal(100.0, 15.0, n)
    f1 = 0.8 * f0 + rng.normal(0.0, 9.0, n)
    f2 = rng.uniform(-100.0, 1000.0, n)
    i0 = rng.integers(-10**6, 2 * 10**6, n)
    cols, arrow = {}, {}
    for name, v in (("f0", f0), ("f1", f1), ("f2", f2), ("i0", i0)):
        valid = rng.random(n) >= null_frac
        if name in all_null:
            valid[:] = False
        bm = S.pack_validity(valid) if null_frac > 0 or name in all_null else None
        cols[name] = (np.ascontiguousarray(v), bm)
        arrow[name] = pa.array(v, mask=~valid if bm is not None else None)
    return cols, pa.table(arrow)


def _close(a, b, rel):
    if a is None or (isinstance(a, float) and math.isnan(a)):
        return b is None or math.isnan(b)
    return abs(a - b) <= rel * max(1.0, abs(b))


@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("n", [0, 1, 2, 63, 64, 65, 1000, 200_003])
@pytest.mark.parametrize("null_frac", [0.0, 0.05])
def test_c_scans_match_the_numpy_oracle(n, null_frac, threads):
    S.lib().to_set_num_threads(threads)
    cols, t = _table(n, seed=n + 1, null_frac=null_frac)
    tc = O.table_cols(t)
    # COUNT(c)
    for name in cols:
        assert S.count_valid(cols[name][1], n) == int(tc[name].valid.sum())
    # MIN / MAX / SUM / AVG / VARIANCE on f64
    mn, mx, c = S.min_max_f64(*cols["f0"])
    assert c == int(tc["f0"].valid.sum())
    if c:
        assert mn == O.stat_value(tc["f0"], "Min") and mx == O.stat_value(tc["f0"], "Max")
    s, c = S.sum_f64(*cols["f1"])
    if c:
        assert _close(s, O.stat_value(tc["f1"], "Sum"), 1e-9) and _close(s / c, O.stat_value(tc["f1"], "Mean"), 1e-9)
    var, c = S.var_f64(*cols["f0"])
    want = O.stat_value(tc["f0"], "Variance")
    assert (want is None and math.isnan(var)) or _close(var, want, 1e-9)
    # Int64: exact min / max / wrapping SUM
    imn, imx, isum, fsum, c = S.min_max_sum_i64(*cols["i0"])
    if c:
        assert float(imn) == O.stat_value(tc["i0"], "Min") and float(imx) == O.stat_value(tc["i0"], "Max")
        assert float(isum) == O.stat_value(tc["i0"], "Sum")
    # CORR / COVAR_SAMP over pairwise-complete rows (one pass)
    r, cov, c = S.corr_f64(cols["f0"][0], cols["f0"][1], cols["f1"][0], cols["f1"][1])
    x, y = O.pair_values(t, "f0", "f1")
    assert c == len(x)
    pr, pc = O.pearson(x, y), O.covar_samp(x, y)
    assert (pr is None and math.isnan(r)) or _close(r, pr, 1e-9)
    assert (pc is None and math.isnan(cov)) or _close(cov, pc, 1e-9)
    # COUNT(CASE WHEN f2 > 0 AND i0 < 1000000 THEN 1 END)
    got = S.pred_gt_lt(cols["f2"][0], cols["f2"][1], 0.0, cols["i0"][0], cols["i0"][1], 1000000)
    sat, total = O.predicate_counts(t, "f2 > 0 AND i0 < 1000000")[:2] if n else (0, 0)
    assert got == sat and total == n


@pytest.mark.parametrize("n", [5, 70_001])
def test_suite_schedules_agree_and_match_the_oracle(n):
    """the per-constraint schedule (one scan each, run_sequential) and the fused one-scan variant report the same
    metrics, and those are the oracle's constraint metrics"""
    S.use_all_host_threads()
    cols, t = _table(n, seed=9)
    a, b = S.numeric_suite(cols, n), S.numeric_suite_fused(cols, n)
    want = {"size": float(n), "min_f0": O.statistic(t, "f0", "Min", ("GreaterThan", -1e300)).metric,
            "mean_f1": O.statistic(t, "f1", "Mean", ("GreaterThan", -1e300)).metric,
            "corr_f0_f1": O.correlation(t, "f0", "f1", "Pearson", ("GreaterThan", -2.0)).metric,
            "satisfies": O.custom_sql(t, "f2 > 0 AND i0 < 1000000").metric}
    for k in want:
        tol = 0.0 if k in ("size", "min_f0", "satisfies") else 1e-9
        assert abs(a[k] - want[k]) <= tol * max(1.0, abs(want[k])), (k, a[k], want[k])
        assert abs(b[k] - want[k]) <= tol * max(1.0, abs(want[k])), (k, b[k], want[k])


def test_all_null_and_bitmapless_columns():
    S.use_all_host_threads()
    n = 10_000
    cols, t = _table(n, seed=4, all_null=("f0",))
    mn, mx, c = S.min_max_f64(*cols["f0"])
    assert c == 0 and mn == math.inf and mx == -math.inf  # MIN over no rows: the SQL NULL (statistics.rs:284-301)
    r, cov, c = S.corr_f64(cols["f0"][0], cols["f0"][1], cols["f1"][0], cols["f1"][1])
    assert c == 0 and math.isnan(r) and math.isnan(cov)
    var, c = S.var_f64(*cols["f0"])
    assert c == 0 and math.isnan(var)
    out = S.numeric_suite_fused(cols, n)
    assert out["min_f0"] == math.inf and math.isnan(out["corr_f0_f1"])
    cols, t = _table(n, seed=5, null_frac=0.0)
    assert all(bm is None for _, bm in cols.values())
    assert S.count_valid(None, n) == n
    a, b = S.numeric_suite(cols, n), S.numeric_suite_fused(cols, n)
    assert a["min_f0"] == b["min_f0"] == float(np.min(cols["f0"][0]))
    assert abs(a["corr_f0_f1"] - float(np.corrcoef(cols["f0"][0], cols["f1"][0])[0, 1])) < 1e-12
