#!/bin/bash
# GPU check of the fused Huffman path: one-tile debug, parity tests, config-1 and Huffman-only config-3 bench lines
timeout 120 python probes/dbg_huff2.py
timeout 600 python -m pytest tests/test_gpu_huffman.py tests/test_gpu_baseline_sizes.py tests/test_gpu_odd_shapes.py tests/test_gpu_mixed_batches.py tests/test_gpu_malformed.py tests/test_gpu_golden_pins.py tests/test_gpu_nulls.py -x -q 2>&1 | tail -8
python bench.py --config 1 --steps 5 --warmup 2 --no-e2e --cpu-seconds 0.3 2>&1 | python probes/bench_line.py
python bench.py --config 3 --codecs GvrsHuffman --steps 5 --warmup 2 --no-e2e --cpu-seconds 0.3 2>&1 | python probes/bench_line.py
