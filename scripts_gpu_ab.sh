#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 600 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | cut -c1-300 | tail -8
for mode in pdl nopdl; do
  if [ $mode = nopdl ]; then export B200_NO_PDL=1; else unset B200_NO_PDL; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  tail -2 gpurun_out/bench_$mode.err | cut -c1-300
  python -c "
import json;d=json.load(open('gpurun_out/bench_$mode.json'))
print('$mode', d['value'],'tok/s ttft',d['ttft_p50_ms'],'decode ms/step',d['decode_ms_per_token_step'],'hbm frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'], 'e2e', d['e2e']['value'])"
done
