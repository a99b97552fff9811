#!/bin/bash
# pytest -m gpu + the default bench line (what the driver runs at round end), no profiler
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/final_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|error" gpurun_out/final_tests.log | tail -3 | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_final.json'))
print('N=1', d['value'],'tok/s e2e',d['e2e']['value'],'ttft',d['ttft_p50_ms'],'ms/step',d['decode_ms_per_token_step'],'frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'],'cpu',d.get('cpu_baseline'),d['clocks'], 'launches', d['gpu_launches'])"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
