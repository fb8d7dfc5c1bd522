#!/bin/bash
for P in 8 4 3 2; do echo "G4_MATCH_CTAS_PER_SM=$P"; G4_MATCH_CTAS_PER_SM=$P python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  encode', d['encode'])
"; done
