#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + top stall-sample SASS lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "sm__inst_executed_pipe_lsu.sum"]
for v in rows[2:]:
    print("== kernel", v[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k} [{units[i]}] = {v[i]}")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                if float(v[i]) > 0.3:
                    print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]} = {float(v[i]):.2f}")
            except ValueError:
                pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isrc, ist, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[ist] or 0), int(r[iex] or 0), r[ia][-5:], r[isrc]))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print(f"total samples {tot}, total warp-instructions {sum(d[1] for d in data)}")
for d in sorted(data, key=lambda x: -x[0])[:topn]:
    print(f"  {100.0 * d[0] / max(tot, 1):5.1f}%  exec={d[1]:>10}  {d[2]}  {d[3][:100]}")
