#!/bin/bash
# sweep of the fused Huffman decoder's first-pass lookback (config-3 shard, Huffman only)
for SB in 448 640; do for LB in 192 256 320; do
  echo "subBits $SB lookback $LB"
  G4_H2_SUBBITS=$SB G4_H2_LOOKBACK=$LB python bench.py --config 3 --codecs GvrsHuffman --steps 5 --warmup 2 --no-e2e --cpu-seconds 0.2 2>&1 | python probes/bench_line.py | cut -c75-140
done; done
