import sys, os, time, torch
sys.path.insert(0, os.getcwd())
import bench as B, term_b200 as T
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ctx = T.SessionContext(0)
for n in (1000, 100_000_000):
    cols, keep = B.make_device_table(torch, n, 44, dev)
    name = f"d{n}"
    ctx.register_device_table(name, {k: {kk: vv for kk, vv in v.items() if kk not in ("tensor", "bits")} for k, v in cols.items()}, keepalive=keep)
    plan, slots = B.build_suite(T, name).build_plan()
    for _ in range(20): plan.execute(ctx, name)
    K = 500
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(K): plan.execute(ctx, name)
    dt = (time.perf_counter() - t0) / K * 1e6
    print(f"rows={n}: {dt:.1f} us per execute, scan_ms {plan.stats()['scan_ms']*1e3:.1f} us")
ctx.close()
