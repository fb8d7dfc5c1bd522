#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_huffman.py tests/test_gpu_huffman_fused.py tests/test_gpu_malformed.py tests/test_gpu_mixed_batches.py -x -q 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:huffman2 -c 4 --csv --log-file gpurun_out/huff_check.csv python bench.py --config 1 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.2 > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/huff_check.csv | tail -2
python bench.py --config 3 --codecs GvrsHuffman --steps 10 --warmup 3 --no-e2e --cpu-seconds 0.3 2>&1 | python probes/bench_line.py
python bench.py --config 1 --steps 10 --warmup 3 --no-e2e --cpu-seconds 0.3 2>&1 | python probes/bench_line.py
