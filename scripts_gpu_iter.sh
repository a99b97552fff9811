#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -k "tcgen05" 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | cut -c1-600 | tail -12
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 600 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | cut -c1-600 | tail -8
for mode in default mma; do
  unset B200_ATTN_MMA
  [ $mode = mma ] && export B200_ATTN_MMA=1
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  tail -1 gpurun_out/bench_$mode.err | cut -c1-300
  python -c "
import json;d=json.load(open('gpurun_out/bench_$mode.json'))
print('$mode', d['value'],'tok/s ttft',d['ttft_p50_ms'],'decode ms/step',d['decode_ms_per_token_step'],'hbm frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'])"
done
