#!/bin/bash
ncu --set full --import-source on --clock-control none -k regex:"deflate_emit_kernel" -c 1 -o gpurun_out/emit2_full -f python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.1 > gpurun_out/emit2_full.log 2>&1
ls -la gpurun_out/emit2_full.ncu-rep
