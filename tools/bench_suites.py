#!/usr/bin/env python3
"""Secondary benchmarks: BASELINE.json configs C3 (string/PII formats), C4 (uniqueness + foreign key) and
C5 (KLL / grouped completeness / Spearman) at the per-GPU shard sizes of SURVEY §8d (1/8 of the config),
one GPU, HBM-resident. Prints one JSON line per workload with the algorithmic GB/s (SURVEY §8d bytes, each
input buffer counted once) against MEASURED_PEAKS.json. bench.py stays the headline (C2); these lines are
the per-kernel rooflines DESIGN.md §6 quotes.

    python tools/bench_suites.py [c3] [c4] [c5] [--scale 1.0] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import term_b200 as T  # noqa: E402
from term_b200 import _ffi as F  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"])
    return 6650.0


def pack_validity(mask):
    """bool tensor (True = valid) -> LSB-first bitmap bytes, padded with 64 slack bytes"""
    n = mask.numel()
    padn = (-n) % 8
    if padn:
        mask = torch.cat([mask, torch.zeros(padn, dtype=torch.bool, device=mask.device)])
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=mask.device)
    packed = (mask.view(-1, 8).to(torch.uint8) * w).sum(dim=1, dtype=torch.int32).to(torch.uint8)
    out = torch.zeros(packed.numel() + 320 - packed.numel() % 64, dtype=torch.uint8, device=mask.device)
    out[: packed.numel()] = packed
    return out


def validity(n, g, dev, frac):
    chunks = []
    for s in range(0, n, 1 << 26):
        e = min(n, s + (1 << 26))
        chunks.append(torch.rand(e - s, generator=g, device=dev) >= frac)
    return pack_validity(torch.cat(chunks))


def make_strings(n, g, dev, null_frac=0.02):
    """C3 column: lengths ~ clipped Poisson(24) in [8, 64]; 70 % email-shaped, 10 % SSN-shaped (11 bytes),
    10 % 16-digit card-shaped, 10 % lowercase noise; ASCII. Returns (offsets i32, bytes u8, validity, cats)."""
    lens = torch.poisson(torch.full((n,), 24.0, device=dev), generator=g).clamp_(8, 64).to(torch.int64)
    cat = torch.rand(n, generator=g, device=dev)
    is_ssn = (cat >= 0.7) & (cat < 0.8)
    is_card = (cat >= 0.8) & (cat < 0.9)
    is_email = cat < 0.7
    lens[is_ssn] = 11
    lens[is_card] = 16
    offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=offs[1:])
    total = int(offs[-1].item())
    data = torch.randint(97, 123, (total + 256,), generator=g, device=dev, dtype=torch.uint8)
    digits = torch.randint(48, 58, (total + 256,), generator=g, device=dev, dtype=torch.uint8)
    # which row each byte belongs to
    row_of = torch.repeat_interleave(torch.arange(n, device=dev), lens)
    dig_row = (is_ssn | is_card)[row_of]
    data[:total][dig_row] = digits[:total][dig_row]
    del digits, dig_row, row_of
    start = offs[:-1]
    # emails: local@domain.com
    e_idx = torch.nonzero(is_email).squeeze(1)
    es, el = start[e_idx], lens[e_idx]
    data[es + el // 2] = 64  # '@'
    data[es + el - 4] = 46   # '.'
    s_idx = torch.nonzero(is_ssn).squeeze(1)
    ss = start[s_idx]
    # SSN area must not be 000/666/9xx: force first digit 1..5
    data[ss] = (data[ss] - 48) % 5 + 49
    data[ss + 3] = 45
    data[ss + 6] = 45
    # group/serial not all zero: force one non-zero digit each
    data[ss + 5] = (data[ss + 5] - 48) % 9 + 49
    data[ss + 10] = (data[ss + 10] - 48) % 9 + 49
    c_idx = torch.nonzero(is_card).squeeze(1)
    data[start[c_idx]] = 52  # '4' (visa-like prefix)
    v = validity(n, g, dev, null_frac)
    return offs.to(torch.int32), data, v, (is_email, is_ssn, is_card), total


def run_plan(plan, ctx, table, steps, key):
    for _ in range(2):
        plan.execute(ctx, table)
    ms, wall = [], []
    for _ in range(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        plan.execute(ctx, table)
        wall.append((time.perf_counter() - t0) * 1e3)
        ms.append(plan.stats()[key])
    return sum(ms) / len(ms), sum(wall) / len(wall), plan.stats()


def report(name, workload, n_rows, alg_bytes, kernel_ms, wall_ms, stats, extra=None):
    gbs = alg_bytes / (kernel_ms / 1e3) / 1e9
    line = {"workload": name, "config": workload, "rows": n_rows, "algorithmic_bytes": alg_bytes, "kernel_ms": kernel_ms,
            "wall_ms": wall_ms, "rows_per_s": n_rows / (wall_ms / 1e3), "achieved_gbs": gbs, "peak_gbs": peak(),
            "frac": gbs / peak(), "launches": int(stats["launches"]), "engine_bytes_scanned": int(stats["bytes_scanned"])}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def bench_c1(ctx, dev, scale, steps):
    """BASELINE.json configs[0]: quickstart suite on a 1 M-row users table (user_id i64, email Utf8), through the public
    API from HOST Arrow data (register + run), and with the table resident."""
    import numpy as np
    import pyarrow as pa
    n = int(1_000_000 * scale)
    ids = np.arange(n, dtype=np.int64)
    emails = pa.array([f"user{i}@example.com" for i in range(n)], type=pa.string())
    table = pa.table({"user_id": pa.array(ids), "email": emails})
    check = (T.Check.builder("quickstart").completeness("user_id", 1.0).validates_uniqueness(["email"], 1.0)
             .validates_regex("email", "@", 1.0).build())
    suite = T.ValidationSuite.builder("c1").table_name("users").check(check).build()
    plan, slots = suite.build_plan()
    ctx.register_table("users", table)
    for _ in range(2):
        plan.execute(ctx, "users")
    walls = []
    for _ in range(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        plan.execute(ctx, "users")
        walls.append((time.perf_counter() - t0) * 1e3)
    st = plan.stats()
    res = {plan.result(s).name: plan.result(s).metric for _, _, s in slots}
    ctx.deregister_table("users")
    e2e = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ctx.register_table("users", table)
        plan.execute(ctx, "users")
        _ = [plan.result(s) for _, _, s in slots]
        ctx.deregister_table("users")
        e2e.append((time.perf_counter() - t0) * 1e3)
    alg = int(st["bytes_scanned"])
    print(json.dumps({"workload": "c1_quickstart", "config": "is_complete(user_id) + is_unique(email) + has_pattern(email, '@') on 1 M rows",
                      "rows": n, "resident_wall_ms": sum(walls) / len(walls), "gpu_ms": st["gpu_ms"], "e2e_wall_ms_from_host_arrow": sum(e2e) / len(e2e),
                      "rows_per_s_resident": n / (sum(walls) / len(walls) / 1e3), "rows_per_s_e2e": n / (sum(e2e) / len(e2e) / 1e3),
                      "algorithmic_bytes": alg, "launches": int(st["launches"]), "results": res}), flush=True)


def bench_c3(ctx, dev, scale, steps):
    n = int(25_000_000 * scale)
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 3)
    offs, data, v, cats, total = make_strings(n, g, dev)
    ctx.register_device_table("pii", {"s": dict(dtype=F.TG_UTF8, n_rows=n, values=data.data_ptr(), offsets=offs.data_ptr(),
                                                validity=v.data_ptr(), n_value_bytes=total)}, keepalive=[offs, data, v])
    check = (T.Check.builder("pii").validates_regex("s", "@", 0.5).validates_email("s", 0.5).contains_ssn("s", 0.05)
             .validates_credit_card("s", 0.5, True).build())
    suite = T.ValidationSuite.builder("c3").table_name("pii").check(check).build()
    plan, slots = suite.build_plan()
    kms, wms, st = run_plan(plan, ctx, "pii", steps, "string_ms")
    res = {plan.result(s).name: plan.result(s).metric for _, _, s in slots}
    alg = 4 * (n + 1) + total + (n + 7) // 8
    exp_email = float((cats[0].sum().item())) / n
    report("c3_string_formats", "regex('@') + email + ssn + credit_card(detect_only) on one Utf8 column, avg 24 B, 2% null",
           n, alg, kms, wms, st, {"results": res, "email_shaped_fraction": exp_email})
    # single-pattern variants for the per-pattern cost
    for nm, build in (("regex_at", lambda b: b.validates_regex("s", "@", 0.5)), ("email", lambda b: b.validates_email("s", 0.5))):
        suite = T.ValidationSuite.builder(nm).table_name("pii").check(build(T.Check.builder(nm)).build()).build()
        plan, _ = suite.build_plan()
        kms, wms, st = run_plan(plan, ctx, "pii", steps, "string_ms")
        report("c3_" + nm, "single pattern", n, alg, kms, wms, st)
    ctx.deregister_table("pii")


def bench_x(ctx, dev, scale, steps):
    """length / data-type constraints on the C3 string column, histogram on a 125 M-row f64 column"""
    n = int(25_000_000 * scale)
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 3)
    offs, data, v, cats, total = make_strings(n, g, dev)
    ctx.register_device_table("pii", {"s": dict(dtype=F.TG_UTF8, n_rows=n, values=data.data_ptr(), offsets=offs.data_ptr(),
                                                validity=v.data_ptr(), n_value_bytes=total)}, keepalive=[offs, data, v])
    alg = 4 * (n + 1) + total + (n + 7) // 8
    check = (T.Check.builder("len").has_min_length("s", 12).has_max_length("s", 40).has_length_between("s", 8, 64).is_not_empty("s").build())
    plan, slots = T.ValidationSuite.builder("x_len").table_name("pii").check(check).build().build_plan()
    kms, wms, st = run_plan(plan, ctx, "pii", steps, "string_ms")
    report("x_length", "min_length + max_length + length_between + not_empty on the C3 string column (one pass)", n, alg, kms, wms, st,
           {"results": {plan.result(s).name: plan.result(s).metric for _, _, s in slots}})
    check = T.Check.builder("dt").constraint(T.DataTypeConstraint("s", T.DataType.Integer, 0.0)).build()
    plan, slots = T.ValidationSuite.builder("x_dt").table_name("pii").check(check).build().build_plan()
    kms, wms, st = run_plan(plan, ctx, "pii", steps, "string_ms")
    report("x_data_type_integer", "DataTypeConstraint(Integer) on the C3 string column", n, alg, kms, wms, st)
    ctx.deregister_table("pii")
    del offs, data, v
    m = int(125_000_000 * scale)
    t = torch.zeros(m + 64, dtype=torch.float64, device=dev)
    t[:m].normal_(100.0, 15.0, generator=g)
    vv = validity(m, g, dev, 0.05)
    ctx.register_device_table("h", {"f0": dict(dtype=F.TG_FLOAT64, n_rows=m, values=t.data_ptr(), validity=vv.data_ptr())}, keepalive=[t, vv])
    plan = T.Plan()
    T.HistogramAnalyzer("f0", 100)._add_to(plan)
    kms, wms, st = run_plan(plan, ctx, "h", steps, "gpu_ms")
    report("x_histogram", "HistogramAnalyzer(f0, 100 buckets): fused scan for min/max/moments + bucket pass (column read twice, like the reference)",
           m, 2 * (8 * m + (m + 7) // 8), kms, wms, st)
    ctx.deregister_table("h")


def bench_pq(ctx, dev, scale, steps):
    """Parquet column chunks (uncompressed, PLAIN, 5 % NULLs) -> HBM, against the same columns registered
    from host Arrow arrays; file bytes come from the page cache. Wall time of register + first use (column_buffers
    waits for the copies and the expand kernel)."""
    import tempfile
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    n = int(20_000_000 * scale)
    rng = np.random.default_rng(42 + 8)
    cols = {}
    for k in range(2):
        cols[f"f{k}"] = pa.array(rng.normal(100.0, 15.0, n), mask=rng.random(n) < 0.05)
        cols[f"i{k}"] = pa.array(rng.integers(-10**6, 10**6, n), mask=rng.random(n) < 0.05)
    t = pa.table(cols)
    path = os.path.join(tempfile.gettempdir(), "tg_bench.parquet")
    pq.write_table(t, path, compression="NONE", use_dictionary=False, row_group_size=n)
    fbytes = os.path.getsize(path)
    open(path, "rb").read()  # page cache

    def timed(fn, name):
        ts = []
        for i in range(steps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            for c in cols:
                ctx.column_buffers(name, c)
            ts.append((time.perf_counter() - t0) * 1e3)
            ctx.deregister_table(name)
        return sum(ts[1:]) / steps
    launches0 = F.lib().tg_engine_launch_count(ctx.handle)
    ms_pq = timed(lambda: ctx.register_parquet("pqb", path), "pqb")
    launches = (F.lib().tg_engine_launch_count(ctx.handle) - launches0) // (steps + 1)
    ms_arrow = timed(lambda: ctx.register_table("pqa", t), "pqa")
    arrow_bytes = sum(8 * n + (n + 7) // 8 for _ in cols)
    print(json.dumps({"workload": "x_parquet_to_hbm", "config": "4 columns (2 f64, 2 i64), 5% NULLs, uncompressed PLAIN, one row group", "rows": n,
                      "file_bytes": fbytes, "arrow_bytes": arrow_bytes, "parquet_ms": ms_pq, "arrow_host_ms": ms_arrow,
                      "parquet_rows_per_s": n / (ms_pq / 1e3), "arrow_rows_per_s": n / (ms_arrow / 1e3),
                      "parquet_gbs_of_file": fbytes / (ms_pq / 1e3) / 1e9, "launches": int(launches)}), flush=True)
    os.remove(path)


def bench_c4(ctx, dev, scale, steps):
    n = int(125_000_000 * scale)
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 4)
    keys = torch.randperm(n, generator=g, device=dev, dtype=torch.int64)
    ndup = max(1, n // 1_000_000)
    keys[torch.randint(0, n, (ndup,), generator=g, device=dev)] = keys[torch.randint(0, n, (ndup,), generator=g, device=dev)]
    pad = torch.zeros(64, dtype=torch.int64, device=dev)
    keys = torch.cat([keys, pad])
    v = validity(n, g, dev, 0.01)
    ctx.register_device_table("keys", {"k": dict(dtype=F.TG_INT64, n_rows=n, values=keys.data_ptr(), validity=v.data_ptr())},
                              keepalive=[keys, v])
    suite = (T.ValidationSuite.builder("c4u").table_name("keys")
             .check(T.Check.builder("u").validates_uniqueness(["k"], 0.9).build()).build())
    plan, slots = suite.build_plan()
    kms, wms, st = run_plan(plan, ctx, "keys", steps, "hash_ms")
    r = plan.result(slots[0][2])
    report("c4_is_unique", "validates_uniqueness on i64 keys (permutation + 1e-6 duplicates, 1% null)", n, 8 * n + (n + 7) // 8,
           kms, wms, st, {"metric": r.metric})
    ctx.deregister_table("keys")
    del keys, v
    # the same check over SPARSE keys (64-bit ids spread over the whole range: no bitmap, the radix-partitioned hash path)
    keys = torch.cat([(torch.randperm(n, generator=g, device=dev, dtype=torch.int64) * 1_000_003) ^ 0x5DEECE66D, pad])
    keys[torch.randint(0, n, (ndup,), generator=g, device=dev)] = keys[torch.randint(0, n, (ndup,), generator=g, device=dev)]
    v = validity(n, g, dev, 0.01)
    ctx.register_device_table("keys", {"k": dict(dtype=F.TG_INT64, n_rows=n, values=keys.data_ptr(), validity=v.data_ptr())},
                              keepalive=[keys, v])
    plan, slots = suite.build_plan()
    kms, wms, st = run_plan(plan, ctx, "keys", steps, "hash_ms")
    r = plan.result(slots[0][2])
    report("c4_is_unique_sparse", "validates_uniqueness on sparse i64 keys (ids spread over 64 bits, 1e-6 duplicates, 1% null)", n, 8 * n + (n + 7) // 8,
           kms, wms, st, {"metric": r.metric})
    ctx.deregister_table("keys")
    del keys, v
    # foreign key: child n rows -> parent n/10 rows
    m = max(1000, n // 10)
    parent = torch.cat([torch.randperm(m, generator=g, device=dev, dtype=torch.int64), pad])
    child = torch.cat([torch.randint(0, int(m * (1 + 1e-4)), (n,), generator=g, device=dev, dtype=torch.int64), pad])
    cv = validity(n, g, dev, 0.01)
    ctx.register_device_table("customers", {"id": dict(dtype=F.TG_INT64, n_rows=m, values=parent.data_ptr(), validity=None)},
                              keepalive=[parent])
    ctx.register_device_table("orders", {"customer_id": dict(dtype=F.TG_INT64, n_rows=n, values=child.data_ptr(),
                                                               validity=cv.data_ptr())}, keepalive=[child, cv])
    suite = (T.ValidationSuite.builder("c4f").table_name("orders")
             .check(T.Check.builder("fk").foreign_key("orders.customer_id", "customers.id").build()).build())
    plan, slots = suite.build_plan()
    kms, wms, st = run_plan(plan, ctx, "orders", steps, "hash_ms")
    r = plan.result(slots[0][2])
    report("c4_foreign_key", "foreign_key orders(n) -> customers(n/10), 1e-4 violation headroom, 1% null child keys", n,
           8 * n + (n + 7) // 8 + 8 * m, kms, wms, st, {"metric": r.metric, "status": r.status.name})
    ctx.deregister_table("orders")
    ctx.deregister_table("customers")


def bench_c5(ctx, dev, scale, steps):
    n = int(125_000_000 * scale)
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 5)
    cols, keep = {}, []
    for k in range(4):
        t = torch.zeros(n + 64, dtype=torch.float64, device=dev)
        if k == 0:
            t[:n].normal_(100.0, 15.0, generator=g)
        elif k == 1:
            t[:n].normal_(0.0, 9.0, generator=g)
            t[:n].add_(cols["f0"]["t"][:n], alpha=0.8)
        elif k == 2:
            t[:n].normal_(0.0, 1.0, generator=g).exp_()
        else:
            t[:n].uniform_(0.0, 1000.0, generator=g)
        v = validity(n, g, dev, 0.05)
        cols[f"f{k}"] = dict(dtype=F.TG_FLOAT64, n_rows=n, values=t.data_ptr(), validity=v.data_ptr(), t=t)
        keep += [t, v]
    group_bytes = {}
    for name, card in (("g0", 16), ("g1", 200)):
        ids = torch.randint(0, card, (n,), generator=g, device=dev)
        # group value "G<id:03d>" (4 bytes each)
        offs = (torch.arange(n + 1, device=dev, dtype=torch.int64) * 4).to(torch.int32)
        data = torch.zeros(4 * n + 256, dtype=torch.uint8, device=dev)
        d4 = data[: 4 * n].view(n, 4)
        d4[:, 0] = 71
        d4[:, 1] = (48 + ids // 100).to(torch.uint8)
        d4[:, 2] = (48 + (ids // 10) % 10).to(torch.uint8)
        d4[:, 3] = (48 + ids % 10).to(torch.uint8)
        cols[name] = dict(dtype=F.TG_UTF8, n_rows=n, values=data.data_ptr(), offsets=offs.data_ptr(), validity=None,
                          n_value_bytes=4 * n)
        keep += [offs, data]
        group_bytes[name] = 4 * (n + 1) + 4 * n
    ctx.register_device_table("wide", {k: {kk: vv for kk, vv in d.items() if kk != "t"} for k, d in cols.items()}, keepalive=keep)

    r = T.AnalysisRunner()
    for k in range(4):
        r.add(T.KllSketchAnalyzer(f"f{k}", 256, (0.5, 0.95, 0.99)))
    plan = T.Plan()
    slots = [a._add_to(plan) for a in r.analyzers]
    kms, wms, st = run_plan(plan, ctx, "wide", steps, "sketch_ms")
    q = plan.analyzer_result(slots[0])
    f0 = cols["f0"]["t"][:n]
    report("c5_kll", "KLL k=256 (p50/p95/p99) on 4 f64 columns, 5% null", n, 4 * (8 * n + (n + 7) // 8), kms, wms, st,
           {"kll_f0": getattr(q, "map", None) or str(q)[:300], "f0_exact_median_ignoring_nulls": float(f0[: min(n, 20_000_000)].median().item())})

    plan = T.Plan()
    for gc in (["g0"], ["g1"], ["g0", "g1"]):
        T.GroupedCompletenessAnalyzer("f0", gc)._add_to(plan)
    kms, wms, st = run_plan(plan, ctx, "wide", steps, "hash_ms")
    alg = group_bytes["g0"] + group_bytes["g1"] + (n + 7) // 8
    report("c5_grouped_completeness", "completeness(f0) grouped by g0 (16), g1 (200), (g0,g1)", n, alg, kms, wms, st)

    plan = T.Plan()
    s = T.CorrelationAnalyzer.spearman("f0", "f1")._add_to(plan)
    kms, wms, st = run_plan(plan, ctx, "wide", max(2, steps // 2), "gpu_ms")
    report("c5_spearman", "Spearman(f0,f1) (min ranks)", n, 2 * (8 * n + (n + 7) // 8), kms, wms, st,
           {"rho": str(plan.analyzer_result(s))[:200]})
    ctx.deregister_table("wide")


def bench_sp(ctx, dev, scale, steps):
    """Spearman alone (two f64 columns, 5 % NULLs) — the profile target of the radix sort"""
    n = int(125_000_000 * scale)
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 5)
    f0 = torch.zeros(n + 64, dtype=torch.float64, device=dev)
    f0[:n].normal_(100.0, 15.0, generator=g)
    f1 = torch.zeros(n + 64, dtype=torch.float64, device=dev)
    f1[:n].normal_(0.0, 9.0, generator=g)
    f1[:n].add_(f0[:n], alpha=0.8)
    v0, v1 = validity(n, g, dev, 0.05), validity(n, g, dev, 0.05)
    ctx.register_device_table("sp", {"f0": dict(dtype=F.TG_FLOAT64, n_rows=n, values=f0.data_ptr(), validity=v0.data_ptr()),
                                     "f1": dict(dtype=F.TG_FLOAT64, n_rows=n, values=f1.data_ptr(), validity=v1.data_ptr())},
                              keepalive=[f0, f1, v0, v1])
    plan = T.Plan()
    s = T.CorrelationAnalyzer.spearman("f0", "f1")._add_to(plan)
    kms, wms, st = run_plan(plan, ctx, "sp", steps, "gpu_ms")
    report("c5_spearman", "Spearman(f0,f1) (min ranks)", n, 2 * (8 * n + (n + 7) // 8), kms, wms, st,
           {"rho": plan.analyzer_result(s).metric_double})
    plan = T.Plan()
    T.KllSketchAnalyzer("f0", 256, (0.5, 0.95, 0.99))._add_to(plan)
    T.KllSketchAnalyzer("f1", 256, (0.5, 0.95, 0.99))._add_to(plan)
    kms, wms, st = run_plan(plan, ctx, "sp", steps, "sketch_ms")
    report("kll_2col", "KLL k=256 on 2 f64 columns", n, 2 * (8 * n + (n + 7) // 8), kms, wms, st)
    ctx.deregister_table("sp")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = T.SessionContext(0)
    for w in a.which:
        {"c1": bench_c1, "c3": bench_c3, "c4": bench_c4, "c5": bench_c5, "x": bench_x, "pq": bench_pq, "sp": bench_sp}[w](ctx, dev, a.scale, a.steps)
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()


# =====================================================================================================================
# Driver-visible suite lines: bench.py calls run_driver_suites() after its headline measurement and embeds the returned
# list as `suites` in its JSON line, so that every BASELINE.json config (C1, C2 full set, C3, C4 incl. the NVLink
# shuffle, C5, and the mixed 8-column full constraint set of north_star) has a measured number at every N the driver runs.
# Per-GPU shard = config / 8; weak scaling: each rank generates its own shard (seed + rank) on the device.
# ms_per_step = device time of K steps on the engine's stream (CUDA events), MAX over ranks; algorithmic bytes per
# SURVEY §8d (each distinct input buffer once), whole job; frac = achieved / (n_gpus x measured HBM peak).
# =====================================================================================================================
def _dist():
    import torch.distributed as dist
    return dist


def _timed(plan, ctx, table, steps, world, dev, stat_key="gpu_ms", warm=2):
    from term_b200 import distributed as D
    dist = _dist()
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    def run():
        if world > 1:
            D.execute_distributed(plan, ctx, table)
        else:
            plan.execute(ctx, table)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        run()
    barrier()
    D.PROFILE = {"shuffle_ms": 0.0, "shuffle_bytes": 0, "exchange_ms": 0.0} if world > 1 else None
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    e0.record(stream)
    for _ in range(steps):
        run()
        kms.append(plan.stats()[stat_key])
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / steps
    prof, D.PROFILE = D.PROFILE, None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = {"ms_per_step": float(t.item()), "kernel_ms": sum(kms) / len(kms), "launches_per_step": (ctx.launch_count() - l0) // steps,
           "steps": steps}
    if prof:
        sb = torch.tensor([float(prof["shuffle_bytes"]), prof["shuffle_ms"], prof["exchange_ms"]], dtype=torch.float64, device=dev)
        smax = sb.clone()
        dist.all_reduce(sb)  # totals over ranks
        dist.all_reduce(smax, op=dist.ReduceOp.MAX)
        out["collectives"] = {"nvlink_shuffle_bytes_per_step_all_ranks": int(sb[0].item() / steps),
                              "shuffle_ms_per_step_max_rank": float(smax[1].item()) / steps,
                              "partial_exchange_ms_per_step_max_rank": float(smax[2].item()) / steps,
                              "note": "shuffle = NCCL all-to-all of hash-partitioned keys (torch CUDA events around the collectives); "
                                      "partial exchange = peer-mailbox / all-gather of the partial states + host merge (wall clock)"}
    return out


def _line(name, workload, n_rows, alg_bytes_per_gpu, world, peak_gbs, timing, extra=None):
    ms, kms = timing["ms_per_step"], timing["kernel_ms"]
    gbs = alg_bytes_per_gpu * world / (ms / 1e3) / 1e9
    line = {"name": name, "workload": workload, "rows_per_gpu": n_rows, "n_gpus": world, "ms_per_step": ms, "kernel_ms_rank0": kms,
            "rows_per_s": n_rows * world / (ms / 1e3), "algorithmic_bytes": alg_bytes_per_gpu * world, "achieved_gbs": gbs,
            "frac": gbs / (world * peak_gbs), "kernel_frac_rank0": (alg_bytes_per_gpu / (kms / 1e3) / 1e9 / peak_gbs) if kms > 0 else None,
            "launches_per_step": timing["launches_per_step"], "steps": timing["steps"]}
    if "collectives" in timing:
        line["collectives"] = timing["collectives"]
    if extra:
        line.update(extra)
    return line


def suite_c1(ctx, dev, world, rank, peak_gbs, steps):
    import numpy as np
    import pyarrow as pa
    n = 1_000_000
    ids = np.arange(rank * n, (rank + 1) * n, dtype=np.int64)
    table = pa.table({"user_id": pa.array(ids), "email": pa.array([f"user{i}@example.com" for i in ids], type=pa.string())})
    check = (T.Check.builder("quickstart").completeness("user_id", 1.0).validates_uniqueness(["email"], 1.0)
             .validates_regex("email", "@", 1.0).build())
    plan, slots = T.ValidationSuite.builder("c1").table_name("users").check(check).build().build_plan()
    ctx.register_table("users", table)
    tm = _timed(plan, ctx, "users", steps, world, dev)
    res = {plan.result(s).name: plan.result(s).metric for _, _, s in slots}
    ctx.deregister_table("users")
    e2e = []
    for _ in range(3):  # from host Arrow data every step (register + run), single process view
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.register_table("users", table)
        if world > 1:
            from term_b200.distributed import execute_distributed
            execute_distributed(plan, ctx, "users")
        else:
            plan.execute(ctx, "users")
        _ = [plan.result(s) for _, _, s in slots]
        ctx.deregister_table("users")
        e2e.append((time.perf_counter() - t0) * 1e3)
    nbytes = sum(b.size for c in table.columns for ch in c.chunks for b in ch.buffers() if b is not None)
    return [_line("c1_quickstart", "C1: is_complete(user_id) + is_unique(email) + has_pattern(email, '@') on a 1 M-row users table per GPU",
                  n, nbytes, world, peak_gbs, tm, {"results": res, "e2e_ms_from_host_arrow": sum(e2e[1:]) / 2, "latency_bound": True})]


def suite_c2full(ctx, dev, world, rank, peak_gbs, steps, c2_table, build_full_suite):
    n = ctx.num_rows(c2_table)
    plan, slots = build_full_suite(T, c2_table).build_plan()
    tm = _timed(plan, ctx, c2_table, steps, world, dev, "scan_ms")
    return [_line("c2_full_numeric_set", "C2 full numeric set: {completeness,min,max,mean,sum,stddev} x 8 cols + 4 correlations (52 constraints)",
                  n, 8 * (8 * n + (n + 7) // 8), world, peak_gbs, tm)]


def suite_c3(ctx, dev, world, rank, peak_gbs, steps):
    n = 25_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 3 + 1000 * rank)
    offs, data, v, cats, total = make_strings(n, g, dev)
    ctx.register_device_table("pii", {"s": dict(dtype=F.TG_UTF8, n_rows=n, values=data.data_ptr(), offsets=offs.data_ptr(),
                                                validity=v.data_ptr(), n_value_bytes=total)}, keepalive=[offs, data, v])
    check = (T.Check.builder("pii").validates_regex("s", "@", 0.5).validates_email("s", 0.5).contains_ssn("s", 0.05)
             .validates_credit_card("s", 0.5, True).build())
    plan, slots = T.ValidationSuite.builder("c3").table_name("pii").check(check).build().build_plan()
    tm = _timed(plan, ctx, "pii", steps, world, dev, "string_ms")
    res = {plan.result(s).name: plan.result(s).metric for _, _, s in slots}
    ctx.deregister_table("pii")
    return [_line("c3_string_formats", "C3: regex('@') + email + ssn + credit_card(detect_only) on one Utf8 column, 25 M rows per GPU, avg 24 B, 2% null",
                  n, 4 * (n + 1) + total + (n + 7) // 8, world, peak_gbs, tm, {"results": res})]


def suite_c4(ctx, dev, world, rank, peak_gbs, steps):
    n = 125_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 4 + 1000 * rank)
    pad = torch.zeros(64, dtype=torch.int64, device=dev)
    # keys: rank r holds a permutation of [r n, (r + 1) n) with 1e-6 duplicates: globally unique up to the duplicates
    keys = torch.randperm(n, generator=g, device=dev, dtype=torch.int64) + rank * n
    ndup = max(1, n // 1_000_000)
    keys[torch.randint(0, n, (ndup,), generator=g, device=dev)] = keys[torch.randint(0, n, (ndup,), generator=g, device=dev)]
    keys = torch.cat([keys, pad])
    v = validity(n, g, dev, 0.01)
    ctx.register_device_table("keys", {"k": dict(dtype=F.TG_INT64, n_rows=n, values=keys.data_ptr(), validity=v.data_ptr())}, keepalive=[keys, v])
    plan, slots = (T.ValidationSuite.builder("c4u").table_name("keys")
                   .check(T.Check.builder("u").validates_uniqueness(["k"], 0.9).build()).build().build_plan())
    tm = _timed(plan, ctx, "keys", steps, world, dev, "hash_ms")
    out = [_line("c4_is_unique", "C4: validates_uniqueness on 125 M i64 keys per GPU (permutation + 1e-6 duplicates, 1% null); N > 1: keys "
                                 "hash-partitioned on the device, NCCL all-to-all over NVLink, per-GPU dedup",
                 n, 8 * n + (n + 7) // 8, world, peak_gbs, tm, {"metric": plan.result(slots[0][2]).metric})]
    ctx.deregister_table("keys")
    del keys, v
    # the same check over SPARSE keys (ids spread over 64 bits: no bitmap; the radix-partitioned hash tables on one GPU, the hash
    # shuffle instead of the range shuffle across GPUs) — the unfriendly case of the same configuration
    keys = torch.cat([((torch.randperm(n, generator=g, device=dev, dtype=torch.int64) + rank * n) * 1_000_003) ^ 0x5DEECE66D, pad])
    keys[torch.randint(0, n, (ndup,), generator=g, device=dev)] = keys[torch.randint(0, n, (ndup,), generator=g, device=dev)]
    v = validity(n, g, dev, 0.01)
    ctx.register_device_table("keys", {"k": dict(dtype=F.TG_INT64, n_rows=n, values=keys.data_ptr(), validity=v.data_ptr())}, keepalive=[keys, v])
    plan, slots = (T.ValidationSuite.builder("c4s").table_name("keys")
                   .check(T.Check.builder("u").validates_uniqueness(["k"], 0.9).build()).build().build_plan())
    tm = _timed(plan, ctx, "keys", steps, world, dev, "hash_ms")
    out.append(_line("c4_is_unique_sparse", "C4 with sparse keys: validates_uniqueness on 125 M i64 ids per GPU spread over 64 bits (1e-6 duplicates, 1% null): "
                                            "hashes radix-sorted into 2^16 buckets + shared-memory de-duplication (hashsort.cu); N > 1: hash shuffle over NVLink",
                     n, 8 * n + (n + 7) // 8, world, peak_gbs, tm, {"metric": plan.result(slots[0][2]).metric}))
    ctx.deregister_table("keys")
    del keys, v
    m = n // 10
    parent = torch.cat([torch.randperm(m, generator=g, device=dev, dtype=torch.int64) + rank * m, pad])
    child = torch.cat([torch.randint(0, int(m * world * (1 + 1e-4)), (n,), generator=g, device=dev, dtype=torch.int64), pad])
    cv = validity(n, g, dev, 0.01)
    ctx.register_device_table("customers", {"id": dict(dtype=F.TG_INT64, n_rows=m, values=parent.data_ptr(), validity=None)}, keepalive=[parent])
    ctx.register_device_table("orders", {"customer_id": dict(dtype=F.TG_INT64, n_rows=n, values=child.data_ptr(), validity=cv.data_ptr())},
                              keepalive=[child, cv])
    plan, slots = (T.ValidationSuite.builder("c4f").table_name("orders")
                   .check(T.Check.builder("fk").foreign_key("orders.customer_id", "customers.id").build()).build().build_plan())
    tm = _timed(plan, ctx, "orders", steps, world, dev, "hash_ms")
    r = plan.result(slots[0][2])
    out.append(_line("c4_foreign_key", "C4: foreign_key orders (125 M per GPU) -> customers (12.5 M per GPU), 1e-4 violation headroom, 1% null "
                                       "child keys; N > 1: both sides hash-shuffled over NVLink",
                     n, 8 * n + (n + 7) // 8 + 8 * m, world, peak_gbs, tm, {"metric": r.metric, "status": r.status.name}))
    ctx.deregister_table("orders")
    ctx.deregister_table("customers")
    return out


def _c5_columns(n, g, dev, n_float=4):
    cols, keep = {}, []
    for k in range(n_float):
        t = torch.zeros(n + 64, dtype=torch.float64, device=dev)
        if k == 0:
            t[:n].normal_(100.0, 15.0, generator=g)
        elif k == 1:
            t[:n].normal_(0.0, 9.0, generator=g)
            t[:n].add_(cols["f0"]["t"][:n], alpha=0.8)
        elif k == 2:
            t[:n].normal_(0.0, 1.0, generator=g).exp_()
        else:
            t[:n].uniform_(0.0, 1000.0, generator=g)
        v = validity(n, g, dev, 0.05)
        cols[f"f{k}"] = dict(dtype=F.TG_FLOAT64, n_rows=n, values=t.data_ptr(), validity=v.data_ptr(), t=t)
        keep += [t, v]
    return cols, keep


def _group_column(n, card, g, dev):
    ids = torch.randint(0, card, (n,), generator=g, device=dev)
    offs = (torch.arange(n + 1, device=dev, dtype=torch.int64) * 4).to(torch.int32)
    data = torch.zeros(4 * n + 256, dtype=torch.uint8, device=dev)
    d4 = data[: 4 * n].view(n, 4)
    d4[:, 0] = 71
    d4[:, 1] = (48 + ids // 100).to(torch.uint8)
    d4[:, 2] = (48 + (ids // 10) % 10).to(torch.uint8)
    d4[:, 3] = (48 + ids % 10).to(torch.uint8)
    return dict(dtype=F.TG_UTF8, n_rows=n, values=data.data_ptr(), offsets=offs.data_ptr(), validity=None, n_value_bytes=4 * n), [offs, data], 4 * (n + 1) + 4 * n


def suite_c5(ctx, dev, world, rank, peak_gbs, steps):
    n = 125_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 5 + 1000 * rank)
    cols, keep = _c5_columns(n, g, dev)
    gb = {}
    for name, card in (("g0", 16), ("g1", 200)):
        cols[name], k2, gb[name] = _group_column(n, card, g, dev)
        keep += k2
    ctx.register_device_table("wide", {k: {kk: vv for kk, vv in d.items() if kk != "t"} for k, d in cols.items()}, keepalive=keep)
    out = []
    plan = T.Plan()
    slots = [T.KllSketchAnalyzer(f"f{k}", 256, (0.5, 0.95, 0.99))._add_to(plan) for k in range(4)]
    tm = _timed(plan, ctx, "wide", steps, world, dev, "sketch_ms")
    q = plan.analyzer_result(slots[0]).map
    out.append(_line("c5_kll", "C5: KLL k=256 (p50/p95/p99) on 4 f64 columns, 125 M rows per GPU, 5% null", n, 4 * (8 * n + (n + 7) // 8),
                     world, peak_gbs, tm, {"kll_f0": q}))
    plan = T.Plan()
    for gc in (["g0"], ["g1"], ["g0", "g1"]):
        T.GroupedCompletenessAnalyzer("f0", gc)._add_to(plan)
    tm = _timed(plan, ctx, "wide", steps, world, dev, "hash_ms")
    out.append(_line("c5_grouped_completeness", "C5: completeness(f0) grouped by g0 (16 values), g1 (200), (g0, g1)", n,
                     gb["g0"] + gb["g1"] + (n + 7) // 8, world, peak_gbs, tm))
    plan = T.Plan()
    s = T.CorrelationAnalyzer.spearman("f0", "f1")._add_to(plan)
    tm = _timed(plan, ctx, "wide", max(2, steps // 2), world, dev, "gpu_ms", warm=1)
    out.append(_line("c5_spearman", "C5: Spearman(f0, f1) (SQL RANK() = minimum ranks) over the pairwise-complete rows", n,
                     2 * (8 * n + (n + 7) // 8), world, peak_gbs, tm, {"rho": plan.analyzer_result(s).metric_double}))
    ctx.deregister_table("wide")
    return out


def suite_mixed(ctx, dev, world, rank, peak_gbs, steps):
    """north_star's target shape: a 1 B-row (125 M per GPU), 8-column table with the full constraint set — numeric scan,
    string patterns, hash uniqueness, KLL and grouped completeness in ONE plan"""
    n = 125_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(42 + 9 + 1000 * rank)
    cols, keep = _c5_columns(n, g, dev, n_float=3)
    alg = 0
    for k in range(2):
        t = torch.zeros(n + 64, dtype=torch.int64, device=dev)
        t[:n].random_(-10**6, 10**6 + 1, generator=g)
        v = validity(n, g, dev, 0.05)
        cols[f"i{k}"] = dict(dtype=F.TG_INT64, n_rows=n, values=t.data_ptr(), validity=v.data_ptr())
        keep += [t, v]
    key = torch.cat([torch.randperm(n, generator=g, device=dev, dtype=torch.int64) + rank * n, torch.zeros(64, dtype=torch.int64, device=dev)])
    kv = validity(n, g, dev, 0.01)
    cols["k"] = dict(dtype=F.TG_INT64, n_rows=n, values=key.data_ptr(), validity=kv.data_ptr())
    keep += [key, kv]
    alg += 6 * (8 * n + (n + 7) // 8)
    offs, data, sv, cats, total = make_strings(n, g, dev)
    cols["s"] = dict(dtype=F.TG_UTF8, n_rows=n, values=data.data_ptr(), offsets=offs.data_ptr(), validity=sv.data_ptr(), n_value_bytes=total)
    keep += [offs, data, sv]
    alg += 4 * (n + 1) + total + (n + 7) // 8
    cols["g"], k2, gbytes = _group_column(n, 200, g, dev)
    keep += k2
    alg += gbytes
    ctx.register_device_table("mixed", {k: {kk: vv for kk, vv in d.items() if kk != "t"} for k, d in cols.items()}, keepalive=keep)
    A = T.Assertion
    cb = T.Check.builder("full").has_size(A.GreaterThan(0.0))
    for c in ("f0", "f1", "f2", "i0", "i1", "k", "s", "g"):
        cb.completeness(c, 0.9)
    for c in ("f0", "f1", "f2"):
        for st in ("Min", "Max", "Mean", "StandardDeviation"):
            cb.statistic(c, T.StatisticType[st], A.GreaterThan(-1e300))
    for c in ("i0", "i1"):
        for st in ("Min", "Max", "Sum"):
            cb.statistic(c, T.StatisticType[st], A.GreaterThan(-1e300))
    cb.has_correlation("f0", "f1", A.GreaterThan(0.5)).satisfies("f2 > 0 AND i0 < 1000000")
    cb.validates_uniqueness(["k"], 0.9).validates_regex("s", "@", 0.5).validates_email("s", 0.5)
    plan, slots = T.ValidationSuite.builder("mixed").table_name("mixed").check(cb.build()).build().build_plan()
    kll = T.KllSketchAnalyzer("f2", 256, (0.5, 0.95, 0.99))._add_to(plan)
    T.GroupedCompletenessAnalyzer("f0", ["g"])._add_to(plan)
    tm = _timed(plan, ctx, "mixed", steps, world, dev, "gpu_ms")
    st = plan.stats()
    failed = [plan.result(s).name for _, _, s in slots if plan.result(s).status.name != "Success"]
    ctx.deregister_table("mixed")
    return [_line("mixed_full_constraint_set",
                  "north_star target shape: 125 M rows x 8 cols per GPU (1 B rows on 8 GPUs; 3 f64, 2 i64, 1 i64 key, 1 Utf8 avg 24 B, 1 Utf8 group), "
                  "33 constraints + KLL + grouped completeness in one plan: completeness x8, min/max/mean/stddev x3, min/max/sum x2, "
                  "correlation, satisfies, is_unique(k), regex + email(s), KLL(f2), completeness(f0) by g",
                  n, alg, world, peak_gbs, tm,
                  {"breakdown_ms_rank0": {k: st[k] for k in ("scan_ms", "string_ms", "hash_ms", "sketch_ms")}, "constraints_not_success": failed,
                   "kll_f2": plan.analyzer_result(kll).map})]


def run_driver_suites(which, ctx, dev, world, rank, peak_gbs, steps, c2_table, build_full_suite):
    out = []
    for w in which:
        try:
            if w == "c1":
                out += suite_c1(ctx, dev, world, rank, peak_gbs, steps)
            elif w == "c2full":
                out += suite_c2full(ctx, dev, world, rank, peak_gbs, steps, c2_table, build_full_suite)
            elif w == "c3":
                out += suite_c3(ctx, dev, world, rank, peak_gbs, steps)
            elif w == "c4":
                out += suite_c4(ctx, dev, world, rank, peak_gbs, steps)
            elif w == "c5":
                out += suite_c5(ctx, dev, world, rank, peak_gbs, steps)
            elif w == "mixed":
                out += suite_mixed(ctx, dev, world, rank, peak_gbs, steps)
        except torch.cuda.OutOfMemoryError as ex:  # a suite that does not fit must not take the headline line with it
            out.append({"name": w, "error": "out of device memory: " + str(ex)[:160]})
        torch.cuda.empty_cache()
    return out
