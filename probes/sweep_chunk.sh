#!/bin/bash
# staged DEFLATE encoder: input bytes per chunk (scratch = 12x) against the L2 size
for MB in 2048 64 16 8 4; do
  echo "G4_STAGED_CHUNK_MB=$MB"
  G4_STAGED_CHUNK_MB=$MB python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  encode', d['encode'])
"
done
