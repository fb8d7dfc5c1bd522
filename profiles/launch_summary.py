#!/usr/bin/env python3
"""Prints kernel / metric / value per launch from an `ncu --csv --log-file` launch list.  Usage: launch_summary.py file.csv [...]"""
import csv
import sys

for f in sys.argv[1:]:
    print(f)
    rows = [r for r in csv.reader(open(f)) if len(r) > 5 and r[0].isdigit()]
    last = None
    for r in rows:
        name = r[4].replace("unnamed>::", "").replace("void ", "")[:44]
        key = (r[0], name)
        if key != last:
            print("  %-46s grid %-14s" % (name, r[8]), end="")
            last = key
        print("  %s=%s" % (r[-3].split(".")[0].replace("smsp__", "").replace("gpu__", "").replace("dram__", ""), r[-1]), end="")
        if r is rows[-1] or (rows[rows.index(r) + 1][0], rows[rows.index(r) + 1][4].replace("unnamed>::", "").replace("void ", "")[:44]) != key:
            print()
