#!/bin/bash
ncu --set full --import-source on --clock-control none -k regex:"deflate_emit_kernel|deflate_sort_kernel|lsop_encode_kernel" -c 3 -o gpurun_out/emit_full -f python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.1 > gpurun_out/emit_full.log 2>&1
ls -la gpurun_out/emit_full.ncu-rep
