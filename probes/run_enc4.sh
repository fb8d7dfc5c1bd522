#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for C in 1 2 3; do python bench.py --config $C --steps 3 --warmup 3 --no-e2e --cpu-seconds 0.2 2>&1 | python probes/bench_line.py; done
