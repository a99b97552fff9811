#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout 600 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | cut -c1-400 | tail -6
LAYERS=4 timeout 300 python tools/timeline.py 2>&1 | sed -n '1,2p;18,26p'
python tools/gemm_bw.py 2>&1 | sed -n '2,3p'
for i in 1 2; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python -c "
import json;d=json.load(open('gpurun_out/bench_iter.json'))
print(d['value'],'tok/s ttft',d['ttft_p50_ms'],'decode ms/step',d['decode_ms_per_token_step'],'hbm frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'])"
done
