#!/usr/bin/env python3
"""Instruction and sample share per source-line range of an .ncu-rep (--page source).
Usage: ncu_regions.py report.ncu-rep file.cuh:lo-hi[:label] ..."""
import csv, io, subprocess, sys
rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    parts = a.split(":")
    lo, hi = parts[1].split("-")
    regions.append((parts[0], int(lo), int(hi), parts[2] if len(parts) > 2 else a))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur, hdr, rows = None, None, []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        d = dict(zip(hdr, r))
        try:
            rows.append((cur, int(r[0]), int(d["Instructions Executed"]), int(d["# Samples"])))
        except ValueError:
            pass
ti, ts = sum(x[2] for x in rows) or 1, sum(x[3] for x in rows) or 1
seen = 0
for f, lo, hi, label in regions:
    i = sum(x[2] for x in rows if x[0] == f and lo <= x[1] <= hi)
    s = sum(x[3] for x in rows if x[0] == f and lo <= x[1] <= hi)
    seen += i
    print("%5.1f%% inst %5.1f%% smp  %s" % (100 * i / ti, 100 * s / ts, label))
print("%5.1f%% inst elsewhere; %d warp instructions in all" % (100 * (ti - seen) / ti, ti))
