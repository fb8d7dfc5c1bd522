#!/bin/bash
# final round-2 evidence after the encode work: ncu of the staged DEFLATE kernels, launch list of one encode pass, bench lines
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"deflate_(sort|match_window|decide_ring|emit)_kernel" -c 4 -o gpurun_out/r02b_deflate_encode -f python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.1 > gpurun_out/r02b_deflate_encode.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02b_launches_config3_encode.csv python bench.py --config 3 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_config3.csv python bench.py --steps 2 --warmup 1 --no-e2e --cpu-seconds 0.3 > /dev/null 2>&1
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_reference.json 2> gpurun_out/r02b_reference.err
python bench.py --config 1 --steps 10 --warmup 3 > gpurun_out/r02b_c1.json 2>/dev/null
python bench.py --config 2 --steps 10 --warmup 3 > gpurun_out/r02b_c2.json 2>/dev/null
python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r02b_c4.json 2>/dev/null
python bench.py --config 3 --codecs GvrsHuffman --steps 10 --warmup 3 > gpurun_out/r02b_c3h.json 2>/dev/null
for f in bench c1 c2 c4 c3h; do python probes/bench_line.py < gpurun_out/r02b_$f.json; done
tail -2 gpurun_out/r02b_bench.err
