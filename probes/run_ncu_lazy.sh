#!/bin/bash
ncu --set full --import-source on --clock-control none -k regex:deflate_lazy -c 1 -o gpurun_out/lazy_full -f python bench.py --config 4 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.1 > gpurun_out/lazy_full.log 2>&1
ls -la gpurun_out/lazy_full.ncu-rep
