#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_huffman.py tests/test_gpu_huffman_fused.py tests/test_gpu_canon_lsop.py tests/test_gpu_deflate_float.py tests/test_gpu_lsop08.py -x -q 2>&1 | tail -5 > gpurun_out/r02_final_memcheck.txt
cat gpurun_out/r02_final_memcheck.txt
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_huffman_fused.py "tests/test_gpu_canon_lsop.py::test_lsop_decode_bit_exact" -x -q 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|hazard" | sort | uniq -c | sort -rn | head -12 > gpurun_out/r02_final_racecheck.txt
cat gpurun_out/r02_final_racecheck.txt
