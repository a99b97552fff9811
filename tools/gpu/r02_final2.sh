#!/bin/bash
# last check of the round: every -m gpu test on the committed build + a short bench line (no CPU leg)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_final2_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/r02_final2_tests.log | tail -2
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final2_bench.json 2> gpurun_out/r02_final2_bench.err
echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r02_final2_bench.json'))
print('N=1', d['value'],'tok/s e2e',d['e2e']['value'],'ttft',d['ttft_p50_ms'],'ms/step',d['decode_ms_per_token_step'],'frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'],d['clocks'], d['parity_check']['ok'])"
