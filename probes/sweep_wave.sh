#!/bin/bash
# wavefront kernel: warps per CTA x CTAs per SM (libg4codec_w<warps>c<ctas>.so built with -DG4_WAVE_WARPS / -DG4_WAVE_CTAS)
M="--metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:lsop3_ -c 2 --csv"
cp gridfour_b200/libg4codec.so /tmp/base.so
for V in base w4c5 w4c6 w2c10; do
  if [ $V = base ]; then cp /tmp/base.so gridfour_b200/libg4codec.so; else cp gridfour_b200/libg4codec_$V.so gridfour_b200/libg4codec.so; fi
  timeout 300 ncu $M --log-file gpurun_out/wave_$V.csv python bench.py --steps 1 --warmup 1 --no-e2e --cpu-seconds 0.2 > /dev/null 2>&1
  echo $V; python profiles/launch_summary.py gpurun_out/wave_$V.csv | tail -1
done
