#!/bin/bash
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_deflate_encode.py tests/test_gpu_deflate_float.py -x -q 2>&1 | tail -5 > gpurun_out/r02b_memcheck.txt
cat gpurun_out/r02b_memcheck.txt
