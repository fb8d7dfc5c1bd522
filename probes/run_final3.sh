#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r02c_c4.json 2>/dev/null
for f in bench c4; do python probes/bench_line.py < gpurun_out/r02c_$f.json; done
tail -2 gpurun_out/r02c_bench.err
