#!/bin/bash
for P in 32 16 8 4; do echo "G4_DECIDE_LANES=$P"; G4_DECIDE_LANES=$P python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  encode', d['encode'])
"; done
