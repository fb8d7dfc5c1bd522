#!/bin/bash
ncu --set full --import-source on --clock-control none -k regex:"inflate|float_finish|float_decode" -c 2 -o gpurun_out/inflate_full -f python bench.py --config 4 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.1 > gpurun_out/inflate_full.log 2>&1
ls -la gpurun_out/inflate_full.ncu-rep
