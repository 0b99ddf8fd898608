import sys, os, tempfile
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, pyarrow as pa
import term_b200 as T
from tests.test_parquet import _write_encoded
ctx = T.SessionContext(0)
tmp = tempfile.mkdtemp()
path, t = _write_encoded(tmp, 200_000, 0.05, "1.0", "SNAPPY", seed=21, dict_limit=8192)
A = T.Assertion
def suite(name):
    cb = (T.Check.builder("c").has_size(A.GreaterThan(0.0)).completeness("cat", 0.9).has_min("runs", A.LessThan(1.0))
          .has_max("wide", A.GreaterThan(0.0)).has_sum("cat", A.LessThan(1e18)).has_mean("runs", A.Between(0.0, 5.0))
          .has_standard_deviation("wide", A.GreaterThan(0.0)).has_correlation("cat", "wide", A.Between(-1.0, 1.0))
          .satisfies("const = 42").validates_uniqueness(["wide"], 0.0).has_approx_quantile("runs", 0.5, A.Between(0.0, 5.0))
          .completeness("suniq", 0.5).validates_email("suniq", 0.1).validates_regex("scat", "é", 0.01).validates_uniqueness(["suniq"], 0.1)
          .has_min_length("sreq", 0).has_max_length("scat", 100))
    return [(r.name, r.metric) for r in T.ValidationSuite.builder("s").table_name(name).check(cb.build()).build().run(ctx).report.results]
# churn the block cache like the other tests do
for n in (1000, 300_000, 70_001):
    p2, t2 = _write_encoded(tmp, n, 0.3, "2.0", "NONE", seed=n)
    ctx.register_parquet("churn", p2); ctx.deregister_table("churn")
for rep in range(8):
    ctx.register_parquet("pq", path)
    ctx.register_table("ar", t)
    a1, a2, b1, b2 = suite("pq"), suite("pq"), suite("ar"), suite("ar")
    d = [(x, y) for x, y in zip(a1, b1) if x != y]
    print(rep, "pq==pq", a1 == a2, "ar==ar", b1 == b2, "pq==ar", a1 == b1, d[:2])
    ctx.deregister_table("pq"); ctx.deregister_table("ar")
