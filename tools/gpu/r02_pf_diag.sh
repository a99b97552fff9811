#!/bin/bash
mkdir -p gpurun_out
for cfg in "off B200_NO_PREFETCHER=1" "cap16 B200_PF_CAP_KB=16" "cap8 B200_PF_CAP_KB=8" "cap32 B200_PF_CAP_KB=32" "cap64 B200_PF_CAP_KB=64"; do
  set -- $cfg; name=$1; shift
  env "$@" LAYERS=8 timeout 300 python tools/timeline.py > gpurun_out/r02_tl_$name.txt 2>&1
  echo "== $name"; head -1 gpurun_out/r02_tl_$name.txt | cut -c1-150; grep -v prefetch gpurun_out/r02_tl_$name.txt | sed -n 12,19p | cut -c1-110
done
grep prefetch gpurun_out/r02_tl_cap16.txt | head -8
