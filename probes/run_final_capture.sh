#!/bin/bash
# final round-2 evidence: full ncu of the LSOP12 decode set, launch list of a bench run, the default bench line
timeout 400 ncu --set full --import-source on --clock-control none -k regex:lsop[23]_ -c 3 -o gpurun_out/r02_final_lsop_full -f python bench.py --steps 1 --warmup 0 --no-e2e --cpu-seconds 0.3 > gpurun_out/r02_final_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --cpu-seconds 0.3 > gpurun_out/r02_final_launches.log 2>&1
python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_reference.json 2> gpurun_out/r02_final_reference.err
tail -c 600 gpurun_out/r02_final_bench.json; echo; tail -c 400 gpurun_out/r02_final_reference.json
