#!/bin/bash
for i in 1 2; do
for cfg in "s1 X=1" "s2 B200_ATTN_SPLITS=2" "s3 B200_ATTN_SPLITS=3"; do
  set -- $cfg; name=$1; shift
  env "$@" LAYERS=8 ROWS=30 timeout 300 python tools/timeline.py > gpurun_out/r02_tl_attn_$name.txt 2>&1
  echo "$name: $(grep 'step span' gpurun_out/r02_tl_attn_$name.txt) | $(grep attn_dec gpurun_out/r02_tl_attn_$name.txt | sed -n 3p | cut -c1-100)"
done
done
