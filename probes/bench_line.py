"""Prints the headline fields of bench.py JSON lines read from stdin (everything else is passed through, shortened)."""
import json
import sys

for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        r = d.get("roofline") or {}
        print(d["config"]["workload"][:72], "| GB/s %.1f ms %.3f frac %.4f kernel_ms %.3f" % (d["value"], d["ms_per_step"], r.get("frac", 0), r.get("kernel_ms", 0)),
              d["config"].get("tile_choice"), "enc", (d.get("encode") or {}).get("value"))
    else:
        print(l.rstrip()[:200])
