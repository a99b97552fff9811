#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 600 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | cut -c1-700 | tail -25
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | cut -c1-300 | tail -4
python -c "
import json
d=json.load(open('gpurun_out/parity_stats.json'))
for k,v in d.items():
    if 'moe' in k: print(k, v)"
