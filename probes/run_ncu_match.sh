#!/bin/bash
ncu --set full --import-source on --clock-control none -k regex:deflate_match_window -c 1 -o gpurun_out/match_full -f python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.1 > gpurun_out/match_full.log 2>&1
ncu -i gpurun_out/match_full.ncu-rep --page raw --csv 2>/dev/null | python profiles/ncu_regions.py 2>/dev/null | head -5
ls -la gpurun_out/match_full.ncu-rep
