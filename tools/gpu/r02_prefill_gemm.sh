#!/bin/bash
for cfg in "base X=1" "cs B200_2CTA_CS=1" "g16 B200_2CTA_GROUP=16" "g16cs B200_2CTA_GROUP=16 B200_2CTA_CS=1" "g32cs B200_2CTA_GROUP=32 B200_2CTA_CS=1" "g64cs B200_2CTA_GROUP=64 B200_2CTA_CS=1" "g4cs B200_2CTA_GROUP=4 B200_2CTA_CS=1"; do
  set -- $cfg; name=$1; shift
  echo "$name: $(env "$@" python tools/prefill_gemm.py 2>&1 | tail -1)"
done
