#!/bin/bash
mkdir -p gpurun_out
for cfg in "off B200_NO_PREFETCHER=1" "pub3 B200_PF_PUB=3" "pub2 B200_PF_PUB=2 B200_PF_SKIP=0" "pub1 B200_PF_PUB=1 B200_PF_SKIP=0" "pub3dry B200_PF_PUB=3 B200_PF_MODE=1" "pub3lead48 B200_PF_PUB=3 B200_PF_LEAD=48"; do
  set -- $cfg; name=$1; shift
  env "$@" LAYERS=8 timeout 300 python tools/timeline.py > gpurun_out/r02_tl_$name.txt 2>&1
  echo "== $name"; head -3 gpurun_out/r02_tl_$name.txt | cut -c1-150; sed -n 12,20p gpurun_out/r02_tl_$name.txt | cut -c1-110
done
