#!/bin/bash
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"deflate_(sort|match_window|decide_ring|emit)_kernel" -c 4 -o gpurun_out/r02d_deflate_encode -f python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.1 > gpurun_out/r02d_deflate_encode.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02d_launches_config3_encode.csv python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.1 > /dev/null 2>&1
ls -la gpurun_out/r02d_*
