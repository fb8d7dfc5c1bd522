#!/usr/bin/env python3
"""Summarises an .ncu-rep: headline metrics per kernel (--page raw) and the hottest source lines (--page source).
Usage: python profiles/ncu_hotspots.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(H, r))
    print("==== %s" % d.get("Kernel Name", "?"))
    for i, h in enumerate(H):
        if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(r[i] or 0) > 0.3):
            print("  %-85s %-16s %s" % (h, U[i], r[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur, hdr, kern, out = None, None, None, {}
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) >= 2 and r[0] == "Function Name":
        kern = r[1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr, r))
        try:
            inst, smp = int(d["Instructions Executed"]), int(d["# Samples"])
        except ValueError:
            continue
        g = lambda k: int(d.get(k, 0) or 0)
        out.setdefault(kern, []).append((inst, smp, cur, r[0], r[1].strip()[:80], g("stall_barrier"), g("stall_short_sb"), g("stall_wait"),
                                         g("stall_long_sb"), g("stall_branch_resolving")))
for kern, o in out.items():
    ti, ts = sum(x[0] for x in o) or 1, sum(x[1] for x in o) or 1
    print("==== %s: %d warp instructions, %d samples" % (kern[:70], ti, ts))
    for x in sorted(o, key=lambda x: -x[1])[:top]:
        print("%5.1f%% inst %5.1f%% smp %s:%s bar=%d ssb=%d wait=%d lsb=%d br=%d | %s" % (100 * x[0] / ti, 100 * x[1] / ts, x[2], x[3], x[5], x[6], x[7],
                                                                                   x[8], x[9], x[4]))
