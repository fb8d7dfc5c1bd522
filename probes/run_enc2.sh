#!/bin/bash
echo "--- float band, old match"; G4_MATCH_WINDOW=0 timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -k config4 2>&1 | tail -2
echo "--- float band, old decide"; G4_DECIDE_RING=0 timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -k config4 2>&1 | tail -2
python probes/enc_gap.py 2>&1 | tail -14
