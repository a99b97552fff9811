#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "gemm_store or residual or test_gemm_swiglu" 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | cut -c1-500 | tail -12
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 500 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | cut -c1-500 | tail -6
for i in 1 2; do python tools/prefill_gemm.py; B200_NO_2CTA=1 python tools/prefill_gemm.py; done 2>&1 | grep libkserve
for mode in default no2cta; do
  unset B200_NO_2CTA
  [ $mode = no2cta ] && export B200_NO_2CTA=1
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
  tail -1 gpurun_out/bench_$mode.err | cut -c1-300
  python -c "
import json;d=json.load(open('gpurun_out/bench_$mode.json'))
print('$mode', d['value'],'tok/s ttft',d['ttft_p50_ms'],'decode ms/step',d['decode_ms_per_token_step'],'hbm frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'])"
done
