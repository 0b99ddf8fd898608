import sys, os, json, torch
sys.path.insert(0, os.getcwd())
import bench as B
import term_b200 as T
n = 100_000_000
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ctx = T.SessionContext(0)
cols, keep = B.make_device_table(torch, n, 44, dev)
ctx.register_device_table("data", {k: {kk: vv for kk, vv in v.items() if kk not in ("tensor", "bits")} for k, v in cols.items()}, keepalive=keep)
A = T.Assertion
cb = T.Check.builder("wide").has_size(A.GreaterThan(0.0))
names = [f"f{k}" for k in range(4)] + [f"i{k}" for k in range(4)]
for c in names:
    cb.completeness(c, 0.9)
    for s in ("Min", "Max", "Mean", "Sum", "StandardDeviation"):
        cb.statistic(c, T.StatisticType[s], A.GreaterThan(-1e300))
for a, b in (("f0", "f1"), ("f2", "f3"), ("i0", "i1"), ("f0", "i2"), ("f1", "f2"), ("f3", "i3"), ("i1", "i2"), ("f0", "f3")):
    cb.has_correlation(a, b, A.GreaterThan(-2.0))
for e in ("f2 > 0 AND i0 < 1000000", "f0 > 50", "i1 >= 0 OR i2 >= 0", "f3 < 5 AND f1 > 0"):
    cb.satisfies(e)
suite = T.ValidationSuite.builder("wide").table_name("data").check(cb.build()).build()
plan, slots = suite.build_plan()
for _ in range(3): plan.execute(ctx, "data")
ks = []
for _ in range(10):
    plan.execute(ctx, "data"); ks.append(plan.stats()["scan_ms"])
st = plan.stats()
print(json.dumps({"workload": "wide numeric set: 8 NUM + 8 PAIR + 4 PRED aggregates", "scan_ms": sum(ks)/len(ks), "launches": st["launches"],
                  "bytes_scanned": st["bytes_scanned"], "alg_gbs": st["bytes_scanned"]/ (sum(ks)/len(ks)/1e3)/1e9}))
ctx.close()
