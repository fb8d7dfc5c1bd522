#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_deflate_float.py tests/test_gpu_golden_pins.py tests/test_gpu_baseline_sizes.py tests/test_gpu_gvrs_file.py -x -q 2>&1 | tail -3
python bench.py --config 4 --steps 3 --warmup 3 --no-e2e --cpu-seconds 0.2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('config 4 encode', d['encode'], 'decode ms', d['ms_per_step'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/enc4_launches.csv python bench.py --config 4 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.1 > gpurun_out/enc4_launches.log 2>&1
python profiles/launch_summary.py gpurun_out/enc4_launches.csv 2>/dev/null | grep -v fill_terrain | head -12
