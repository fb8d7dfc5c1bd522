#!/bin/bash
# final round-2 evidence: full ncu of the LSOP12 decode set and the fused Huffman kernels, launch lists, bench lines
timeout 400 ncu --set full --import-source on --clock-control none -k regex:lsop[23]_ -c 3 -o gpurun_out/r02_final_lsop_full -f python bench.py --steps 1 --warmup 0 --no-e2e --cpu-seconds 0.3 > gpurun_out/r02_final_full.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:huffman2 -c 2 -o gpurun_out/r02_final_huff2_full -f python bench.py --config 3 --codecs GvrsHuffman --steps 1 --warmup 0 --no-e2e --cpu-seconds 0.3 > gpurun_out/r02_final_huff2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --cpu-seconds 0.3 > gpurun_out/r02_final_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_final_launches_huffman.csv python bench.py --config 3 --codecs GvrsHuffman --steps 2 --warmup 1 --no-e2e --cpu-seconds 0.3 > /dev/null 2>&1
python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_reference.json 2> gpurun_out/r02_final_reference.err
python bench.py --config 1 --steps 10 --warmup 3 > gpurun_out/r02_final_c1.json 2>/dev/null
python bench.py --config 2 --steps 10 --warmup 3 > gpurun_out/r02_final_c2.json 2>/dev/null
python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r02_final_c4.json 2>/dev/null
python bench.py --config 3 --codecs GvrsHuffman --steps 10 --warmup 3 > gpurun_out/r02_final_c3h.json 2>/dev/null
python bench.py --config 5 --steps 5 --warmup 3 --no-e2e --cpu-seconds 0.2 > gpurun_out/r02_final_sweep_n1.json 2>/dev/null
for f in bench c1 c2 c4 c3h; do python probes/bench_line.py < gpurun_out/r02_final_$f.json; done
