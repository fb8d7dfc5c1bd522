#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_deflate_encode.py tests/test_gpu_deflate_float.py "tests/test_gpu_canon_lsop.py::test_lsop_deflate_alternative_matches_oracle" -x -q 2>&1 | tail -5 > gpurun_out/r02b_memcheck.txt
cat gpurun_out/r02b_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_deflate_encode.py -x -q 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|hazard|Race reported" | sort | uniq -c | sort -rn | head -12 > gpurun_out/r02b_racecheck.txt
cat gpurun_out/r02b_racecheck.txt
