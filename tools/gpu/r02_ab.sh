#!/bin/bash
# same-box A/B of two library builds: kserve_b200/lib/libkserve_b200_old.so vs the current one, alternating
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity-check"
for i in 1 2; do
  for v in old new; do
    if [ $v = old ]; then export B200_LIB_PATH=$PWD/kserve_b200/lib/libkserve_b200_old.so; else unset B200_LIB_PATH; fi
    timeout 600 $B > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
    python -c "
import json;d=json.load(open('gpurun_out/ab_$v.json'))
print('$v', d['value'],'tok/s ttft',d['ttft_p50_ms'],'ms/step',d['decode_ms_per_token_step'],'prefill frac',d['roofline_prefill']['frac'],d['clocks']['sm_mhz'])" || tail -3 gpurun_out/ab_$v.err
  done
done
