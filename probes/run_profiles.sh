#!/bin/bash
# bench lines of the other configs and the fused Huffman profile for profiles/
python bench.py --config 1 --steps 10 --warmup 3 > gpurun_out/r02z_c1.json 2>gpurun_out/r02z_c1.err
python bench.py --config 2 --steps 10 --warmup 3 > gpurun_out/r02z_c2.json 2>gpurun_out/r02z_c2.err
python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r02z_c4.json 2>gpurun_out/r02z_c4.err
python bench.py --config 3 --codecs GvrsHuffman --steps 10 --warmup 3 > gpurun_out/r02z_c3h.json 2>gpurun_out/r02z_c3h.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:huffman2 -c 2 -o gpurun_out/r02z_huff2 -f python bench.py --config 3 --codecs GvrsHuffman --steps 1 --warmup 0 --no-e2e --cpu-seconds 0.3 > gpurun_out/r02z_huff2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02z_launches_huffman.csv python bench.py --config 3 --codecs GvrsHuffman --steps 2 --warmup 1 --no-e2e --cpu-seconds 0.3 > /dev/null 2>&1
ls -la gpurun_out/r02z*
