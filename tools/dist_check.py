#!/usr/bin/env python3
"""Multi-GPU parity + timing check (run under torchrun, one rank per GPU, NCCL):
every rank holds a row shard; execute_distributed (row-sharded partials + hash shuffle of uniqueness / foreign-key
keys over NCCL all-to-all) must reproduce, bit for bit, what ONE GPU computes on the concatenated tables.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tools/dist_check.py [--rows 20000000]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import term_b200 as T  # noqa: E402
from term_b200 import _ffi as F  # noqa: E402
from term_b200.distributed import execute_distributed  # noqa: E402


def bitmap(mask):
    n = mask.numel()
    padn = (-n) % 8
    if padn:
        mask = torch.cat([mask, torch.zeros(padn, dtype=torch.bool, device=mask.device)])
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=mask.device)
    packed = (mask.view(-1, 8).to(torch.uint8) * w).sum(dim=1, dtype=torch.int32).to(torch.uint8)
    out = torch.zeros(packed.numel() + 320 - packed.numel() % 64, dtype=torch.uint8, device=mask.device)
    out[: packed.numel()] = packed
    return out


def pad(t):
    return torch.cat([t, torch.zeros(64, dtype=t.dtype, device=t.device)])


def register(ctx, name, cols):
    spec, keep = {}, []
    for c, (vals, valid) in cols.items():
        v = pad(vals)
        b = bitmap(valid) if valid is not None else None
        keep += [v, b]
        dt = {torch.float64: F.TG_FLOAT64, torch.int64: F.TG_INT64, torch.int32: F.TG_INT32, torch.float32: F.TG_FLOAT32}[vals.dtype]
        spec[c] = dict(dtype=dt, n_rows=vals.numel(), values=v.data_ptr(),
                       validity=b.data_ptr() if b is not None else None)
    torch.cuda.synchronize()  # the engine runs on its own stream: the tensors must be complete before it reads them
    ctx.register_device_table(name, spec, keepalive=keep)


def register_strings(ctx, name, col, ids, valid, width=10):
    """a Utf8 column "k" + zero-padded decimal of ids (width digits, 1 + width bytes per row), built on the device"""
    n = ids.numel()
    dev = ids.device
    digits = torch.empty((n, width + 1), dtype=torch.uint8, device=dev)
    digits[:, 0] = ord("k")
    rem = ids.clone()
    for d in range(width, 0, -1):
        digits[:, d] = (rem % 10 + ord("0")).to(torch.uint8)
        rem //= 10
    data = torch.cat([digits.reshape(-1), torch.zeros(256, dtype=torch.uint8, device=dev)])
    offs = torch.cat([torch.arange(n + 1, dtype=torch.int32, device=dev) * (width + 1), torch.zeros(64, dtype=torch.int32, device=dev)])
    b = bitmap(valid) if valid is not None else None
    torch.cuda.synchronize()  # the engine runs on its own stream: the tensors must be complete before it reads them
    ctx.register_device_table(name, {col: dict(dtype=F.TG_UTF8, n_rows=n, values=data.data_ptr(), offsets=offs.data_ptr(),
                                               validity=b.data_ptr() if b is not None else None, n_value_bytes=n * (width + 1))},
                              keepalive=[data, offs, b])


def gather(t, world):
    sizes = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.numel()], dtype=torch.int64, device=t.device))
    cap = int(max(s.item() for s in sizes))
    buf = torch.zeros(cap, dtype=t.dtype, device=t.device)
    buf[: t.numel()] = t
    outs = [torch.zeros(cap, dtype=t.dtype, device=t.device) for _ in range(world)]
    dist.all_gather(outs, buf)
    return torch.cat([o[: int(s.item())] for o, s in zip(outs, sizes)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20_000_000, help="child / key rows per GPU")
    ap.add_argument("--sparse", action="store_true", help="sparse keys (radix-partitioned hash path instead of bitmaps)")
    a = ap.parse_args()
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = T.SessionContext(local)
    g = torch.Generator(device=dev)
    g.manual_seed(100 + rank)
    n, m = a.rows, max(1000, a.rows // 10)
    mul = 1_000_003 if a.sparse else 1
    total_parents = m * world
    # global parent ids = 0..total_parents-1 (x mul), sharded contiguously; children draw from a slightly larger range
    parent = (torch.arange(rank * m, (rank + 1) * m, device=dev, dtype=torch.int64)[torch.randperm(m, generator=g, device=dev)]) * mul
    child = torch.randint(0, int(total_parents * 1.0001) + 1, (n,), generator=g, device=dev, dtype=torch.int64) * mul
    child_valid = torch.rand(n, generator=g, device=dev) >= 0.01
    keys = (torch.randperm(n, generator=g, device=dev, dtype=torch.int64) + rank * n) * mul
    dup = torch.randint(0, n, (max(1, n // 1000),), generator=g, device=dev)
    keys[dup] = (torch.randint(0, n * world, (dup.numel(),), generator=g, device=dev, dtype=torch.int64)) * mul  # cross-rank duplicates
    keys_valid = torch.rand(n, generator=g, device=dev) >= 0.01
    x = torch.empty(n, dtype=torch.float64, device=dev).normal_(100.0, 15.0, generator=g)
    x_valid = torch.rand(n, generator=g, device=dev) >= 0.05
    # Int64 column correlated with amount (Spearman ~ 0.9), its own NULLs
    score = (x * 3.0 + torch.empty(n, dtype=torch.float64, device=dev).normal_(0.0, 20.0, generator=g)).to(torch.int64)
    score_valid = torch.rand(n, generator=g, device=dev) >= 0.03
    # an Int32 key column (travels through the shuffle as its exactly widened values) and a Float32 measure
    region = (keys % 100_003).to(torch.int32)
    price = x.to(torch.float32)
    register(ctx, "orders", {"customer_id": (child, child_valid), "order_key": (keys, keys_valid), "amount": (x, x_valid),
                             "score": (score, score_valid), "region": (region, keys_valid), "price": (price, x_valid)})
    register(ctx, "customers", {"id": (parent, None)})

    A = T.Assertion
    check = (T.Check.builder("integrity").has_size(A.GreaterThan(0.0)).has_mean("amount", A.Between(90.0, 110.0))
             .validates_uniqueness(["order_key"], 0.9)
             .validates_uniqueness(["order_key", "customer_id"], 0.5)  # composite key: fingerprint shuffle
             .foreign_key("orders.customer_id", "customers.id")
             .validates_uniqueness(["region"], 0.0).has_standard_deviation("price", A.GreaterThan(0.0)).build())
    suite = T.ValidationSuite.builder("dist").table_name("orders").check(check).build()
    plan, slots = suite.build_plan()
    extra = T.UniquenessConstraint(["order_key"], T.UniquenessType.UniqueValueRatio, assertion=A.GreaterThan(0.0))._add_to(plan)
    for _ in range(2):
        execute_distributed(plan, ctx, "orders")
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        execute_distributed(plan, ctx, "orders")
    dist.barrier()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    got = [plan.result(s) for _, _, s in slots] + [plan.result(extra)]
    got = [(r.name, r.status.name, r.metric, (r.message or "").split("Examples")[0]) for r in got]
    # Spearman: global ranks -> the pairwise-complete rows are gathered to one GPU (distributed._gather_spearman_columns)
    sp = T.Plan()
    sp_slot = T.CorrelationAnalyzer.spearman("amount", "score")._add_to(sp)
    execute_distributed(sp, ctx, "orders")
    sp_got = sp.analyzer_result(sp_slot)
    sp_all = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(sp_all, torch.tensor([float(sp_got.u[0]), sp_got.metric_double], dtype=torch.float64, device=dev))
    sp_agree = all(bool((v == sp_all[0]).all().item()) for v in sp_all)

    # scan-only plan: the fused step (tg_plan_execute_exchange: states assembled on the device, one synchronisation), and the
    # same plan through the two-call path (TG_NO_FUSED_EXCHANGE) — both must give the single-GPU answer
    def scan_suite(table):
        cb = (T.Check.builder("scan").has_size(A.GreaterThan(0.0)).completeness("amount", 0.9).completeness("customer_id", 0.9)
              .has_min("amount", A.GreaterThan(-1e9)).has_max("amount", A.LessThan(1e9)).has_mean("amount", A.Between(90.0, 110.0))
              .has_standard_deviation("amount", A.LessThan(100.0)).has_sum("score", A.LessThan(1e300)).has_min("score", A.LessThan(1e18))
              .has_correlation("amount", "score", A.GreaterThan(0.5)).satisfies("amount > 50 AND score < 1000"))
        return T.ValidationSuite.builder("scan").table_name(table).check(cb.build()).build().build_plan()
    sc_plan, sc_slots = scan_suite("orders")
    execute_distributed(sc_plan, ctx, "orders")
    sc_fused = [(r.name, r.status.name, r.metric) for r in (sc_plan.result(s) for _, _, s in sc_slots)]
    os.environ["TG_NO_FUSED_EXCHANGE"] = "1"
    execute_distributed(sc_plan, ctx, "orders")
    del os.environ["TG_NO_FUSED_EXCHANGE"]
    sc_two = [(r.name, r.status.name, r.metric) for r in (sc_plan.result(s) for _, _, s in sc_slots)]
    # variable-size partials through the mailboxes: KLL sketch, grouped completeness, two-phase histogram
    an_plan = T.Plan()
    kll_slot = T.KllSketchAnalyzer("amount", 256, (0.5, 0.95, 0.99))._add_to(an_plan)
    hist_slot = T.HistogramAnalyzer("amount", 20)._add_to(an_plan)
    execute_distributed(an_plan, ctx, "orders")
    kll_got, hist_got = an_plan.analyzer_result(kll_slot), an_plan.analyzer_result(hist_slot)

    # foreign key over Utf8 columns: both sides travel as fingerprint records (counts exact, no examples across GPUs)
    ns = min(n, 2_000_000)
    register_strings(ctx, "s_orders", "cid", child[:ns], child_valid[:ns])
    register_strings(ctx, "s_customers", "id", parent, None)
    fk_check = T.Check.builder("sfk").foreign_key("s_orders.cid", "s_customers.id").build()
    fk_plan, fk_slots = T.ValidationSuite.builder("sfk").table_name("s_orders").check(fk_check).build().build_plan()
    execute_distributed(fk_plan, ctx, "s_orders")
    fk_got = [(r.status.name, r.metric, (r.message or "").split(". Examples")[0]) for r in (fk_plan.result(s) for _, _, s in fk_slots)]

    # reference: everything on rank 0's GPU
    full = {"customer_id": gather(child, world), "cv": gather(child_valid.to(torch.uint8), world).bool(),
            "order_key": gather(keys, world), "kv": gather(keys_valid.to(torch.uint8), world).bool(),
            "amount": gather(x, world), "av": gather(x_valid.to(torch.uint8), world).bool(), "parent": gather(parent, world),
            "score": gather(score, world), "sv": gather(score_valid.to(torch.uint8), world).bool(),
            "region": gather(region, world), "price": gather(price, world)}
    gather_done = {"child": gather(child[:ns].contiguous(), world), "cv": gather(child_valid[:ns].to(torch.uint8).contiguous(), world).bool()}
    ok = True
    if rank == 0:
        register(ctx, "orders_all", {"customer_id": (full["customer_id"], full["cv"]), "order_key": (full["order_key"], full["kv"]),
                                     "amount": (full["amount"], full["av"]), "score": (full["score"], full["sv"]),
                                     "region": (full["region"], full["kv"]), "price": (full["price"], full["av"])})
        register(ctx, "customers_all", {"id": (full["parent"], None)})
        check1 = (T.Check.builder("integrity").has_size(A.GreaterThan(0.0)).has_mean("amount", A.Between(90.0, 110.0))
                  .validates_uniqueness(["order_key"], 0.9)
                  .validates_uniqueness(["order_key", "customer_id"], 0.5)
                  .foreign_key("orders_all.customer_id", "customers_all.id")
                  .validates_uniqueness(["region"], 0.0).has_standard_deviation("price", A.GreaterThan(0.0)).build())
        s1 = T.ValidationSuite.builder("single").table_name("orders_all").check(check1).build()
        p1, sl1 = s1.build_plan()
        e1 = T.UniquenessConstraint(["order_key"], T.UniquenessType.UniqueValueRatio, assertion=A.GreaterThan(0.0))._add_to(p1)
        p1.execute(ctx, "orders_all")
        want = [p1.result(s) for _, _, s in sl1] + [p1.result(e1)]
        want = [(r.name, r.status.name, r.metric, (r.message or "").split("Examples")[0].replace("orders_all", "orders").replace("customers_all", "customers"))
                for r in want]
        for gg, ww in zip(got, want):
            exact = gg[0] not in ("mean", "standard_deviation")
            same = gg[1] == ww[1] and gg[3] == ww[3] and (gg[2] == ww[2] if exact else abs(gg[2] - ww[2]) <= 1e-9 * abs(ww[2]))
            ok = ok and same
            if not same:
                print("MISMATCH", gg, ww, flush=True)
        sp_want = T.CorrelationAnalyzer.spearman("amount", "score").compute(ctx, "orders_all")
        sp_ok = (sp_agree and sp_got.u[0] == sp_want.u[0] and sp_got.error == 0
                 and abs(sp_got.metric_double - sp_want.metric_double) <= 1e-9)
        if not sp_ok:
            print("SPEARMAN MISMATCH", sp_got.u[0], sp_got.metric_double, sp_want.u[0], sp_want.metric_double, sp_agree, flush=True)
        ok = ok and sp_ok
        p2, sl2 = scan_suite("orders_all")
        p2.execute(ctx, "orders_all")
        sc_want = [(r.name, r.status.name, r.metric) for r in (p2.result(s) for _, _, s in sl2)]
        exact_names = {"size", "completeness", "min", "max", "sum", "custom_sql"}
        for label, res in (("fused", sc_fused), ("two-call", sc_two)):
            for gg, ww in zip(res, sc_want):
                same = gg[:2] == ww[:2] and (gg[2] == ww[2] if gg[0] in exact_names else abs(gg[2] - ww[2]) <= 1e-9 * max(1.0, abs(ww[2])))
                ok = ok and same
                if not same:
                    print("SCAN MISMATCH", label, gg, ww, flush=True)
        a2 = T.Plan()
        k2, h2 = T.KllSketchAnalyzer("amount", 256, (0.5, 0.95, 0.99))._add_to(a2), T.HistogramAnalyzer("amount", 20)._add_to(a2)
        a2.execute(ctx, "orders_all")
        kw, hw = a2.analyzer_result(k2), a2.analyzer_result(h2)
        an_ok = (kll_got.error == 0 and hist_got.error == 0 and kll_got.u[0] == kw.u[0] and kll_got.map["min"] == kw.map["min"]
                 and kll_got.map["max"] == kw.map["max"])
        srt = torch.sort(full["amount"][full["av"]]).values
        for q in (0.5, 0.95, 0.99):
            est = kll_got.map["quantile_" + str(q)]
            pos = int(torch.searchsorted(srt, torch.tensor([est], dtype=torch.float64, device=dev)).item())
            an_ok = an_ok and abs(pos / srt.numel() - q) <= 0.01  # far inside 1.65 / sqrt(256)
        an_ok = an_ok and all(hist_got.map[k] == v for k, v in hw.map.items() if k.endswith(".count") or k.endswith(".lower") or k in ("min", "max", "total_count"))
        if not an_ok:
            print("ANALYZER MISMATCH", kll_got.map, kw.map, flush=True)
        ok = ok and an_ok
        register_strings(ctx, "s_orders_all", "cid", gather_done["child"], gather_done["cv"])
        register_strings(ctx, "s_customers_all", "id", full["parent"], None)
        fk1 = T.Check.builder("sfk").foreign_key("s_orders_all.cid", "s_customers_all.id").build()
        fp1, fs1 = T.ValidationSuite.builder("sfk").table_name("s_orders_all").check(fk1).build().build_plan()
        fp1.execute(ctx, "s_orders_all")
        fk_want = [(r.status.name, r.metric, (r.message or "").split(". Examples")[0].replace("s_orders_all", "s_orders").replace("s_customers_all", "s_customers"))
                   for r in (fp1.result(s) for _, _, s in fs1)]
        if fk_got != fk_want:
            exp = int((~torch.isin(gather_done["child"], full["parent"]) & gather_done["cv"]).sum().item()) + int((~gather_done["cv"]).sum().item())
            print("UTF8 FK MISMATCH", fk_got, fk_want, "torch says", exp, "child", gather_done["child"].numel(), "parents", full["parent"].numel(),
                  ctx.num_rows("s_orders_all"), ctx.num_rows("s_customers_all"), flush=True)
            ok = False
        print(json.dumps({"check": "multi_gpu_parity", "utf8_foreign_key": fk_got, "scan_fused": sc_fused, "kll_distributed": kll_got.map, "world": world, "rows_per_gpu": n, "parents_per_gpu": m, "sparse_keys": a.sparse,
                          "ok": ok, "ms_per_execute": ms, "results": got,
                          "spearman": {"pairs": sp_got.u[0], "rho_distributed": sp_got.metric_double, "rho_single_gpu": sp_want.metric_double,
                                       "ranks_agree": sp_agree}}), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
