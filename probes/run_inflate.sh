#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_deflate_float.py tests/test_gpu_golden_pins.py tests/test_gpu_baseline_sizes.py tests/test_gpu_gvrs_file.py tests/test_gpu_malformed.py tests/test_gpu_canon_lsop.py tests/test_gpu_lsop08.py -x -q 2>&1 | tail -3
for C in 4; do python bench.py --config $C --steps 5 --warmup 3 --no-e2e --cpu-seconds 0.2 2>&1 | python probes/bench_line.py; done
