"""time of general (UNIT_PRED) predicates over 50 M resident rows: A/B of evaluator changes"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import term_b200 as T
from term_b200 import _ffi as F
dev = torch.device("cuda", 0)
n = 50_000_000
g = torch.Generator(device=dev); g.manual_seed(1)
f = torch.randn(n + 64, generator=g, device=dev, dtype=torch.float64) * 10 + 5
i = torch.randint(-1000, 1000, (n + 64,), generator=g, device=dev, dtype=torch.int64)
j = torch.randint(1, 50, (n + 64,), generator=g, device=dev, dtype=torch.int64)
ctx = T.SessionContext(0)
ctx.register_device_table("p", {"f": dict(dtype=F.TG_FLOAT64, n_rows=n, values=f.data_ptr()), "i": dict(dtype=F.TG_INT64, n_rows=n, values=i.data_ptr()),
                                "j": dict(dtype=F.TG_INT64, n_rows=n, values=j.data_ptr())}, keepalive=[f, i, j])
for expr in ["f + i > 3 OR i % 7 = 0", "abs(i) * 2 - 5 < f * 1000 AND NOT (i / j >= 10)", "f > 0 AND i < 500"]:
    cb = T.Check.builder("c").satisfies(expr)
    plan, slots = T.ValidationSuite.builder("s").table_name("p").check(cb.build()).build().build_plan()
    for _ in range(3):
        plan.execute(ctx, "p")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        plan.execute(ctx, "p")
    torch.cuda.synchronize()
    print(os.environ.get("TG_LIB", "default").split("/")[-1], expr, round((time.perf_counter() - t0) / 20 * 1e3, 4), "ms", plan.result(slots[0][2]).metric)
