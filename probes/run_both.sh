#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_huffman.py tests/test_gpu_huffman_fused.py tests/test_gpu_canon_lsop.py tests/test_gpu_baseline_sizes.py tests/test_gpu_odd_shapes.py tests/test_gpu_mixed_batches.py tests/test_gpu_malformed.py tests/test_gpu_golden_pins.py -x -q 2>&1 | tail -3
M="--metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:lsop[23]_ -c 3 --csv"
timeout 300 ncu $M --log-file gpurun_out/lsop_check.csv python bench.py --steps 1 --warmup 0 --no-e2e --cpu-seconds 0.3 > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/lsop_check.csv | tail -3
python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0.3 2>&1 | python probes/bench_line.py
python bench.py --config 3 --codecs GvrsHuffman --steps 10 --warmup 3 --no-e2e --cpu-seconds 0.3 2>&1 | python probes/bench_line.py
python bench.py --config 1 --steps 10 --warmup 3 --no-e2e --cpu-seconds 0.3 2>&1 | python probes/bench_line.py
