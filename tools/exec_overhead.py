import sys, os, time, torch
sys.path.insert(0, os.getcwd())
import bench as B, term_b200 as T
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ctx = T.SessionContext(0)
n = 1000
cols, keep = B.make_device_table(torch, n, 44, dev)
ctx.register_device_table("d", {k: {kk: vv for kk, vv in v.items() if kk not in ("tensor", "bits")} for k, v in cols.items()}, keepalive=keep)
A = T.Assertion
def timeit(plan, K=2000):
    for _ in range(50): plan.execute(ctx, "d")
    t0 = time.perf_counter()
    for _ in range(K): plan.execute(ctx, "d")
    return (time.perf_counter() - t0) / K * 1e6
p0, _ = T.ValidationSuite.builder("s").table_name("d").check(T.Check.builder("c").has_size(A.GreaterThan(0.0)).build()).build().build_plan()
print("size only (no kernel): %.1f us" % timeit(p0))
p1, _ = T.ValidationSuite.builder("s").table_name("d").check(T.Check.builder("c").has_min("f0", A.GreaterThan(-1e9)).build()).build().build_plan()
print("one NUM aggregate: %.1f us (scan_ms %.1f us)" % (timeit(p1), p1.stats()["scan_ms"] * 1e3))
p5, _ = B.build_suite(T, "d").build_plan()
print("literal suite: %.1f us (scan_ms %.1f us)" % (timeit(p5), p5.stats()["scan_ms"] * 1e3))
ctx.close()
