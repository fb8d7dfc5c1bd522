#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_canon_lsop.py tests/test_gpu_baseline_sizes.py tests/test_gpu_odd_shapes.py tests/test_gpu_mixed_batches.py tests/test_gpu_malformed.py tests/test_gpu_golden_pins.py -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0.3 2>&1 | python probes/bench_line.py
python bench.py --config 2 --steps 10 --warmup 3 --no-e2e --cpu-seconds 0.3 2>&1 | python probes/bench_line.py
python bench.py --config 5 --steps 5 --warmup 3 --no-e2e --cpu-seconds 0.2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for s in d['sweep']: print(s['tile'], round(s['decode_gbs'],1), round(s['ms_per_step'],3), round(s['hbm_frac_per_gpu'],4))
"
