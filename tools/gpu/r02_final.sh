#!/bin/bash
# final round-2 check: every -m gpu test, smoke(), the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02_final_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/r02_final_tests.log | tail -2
cp gpurun_out/parity_stats.json gpurun_out/r02_final_parity_stats.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r02_final_bench.json'))
print('N=1', d['value'],'tok/s e2e',d['e2e']['value'],'ttft',d['ttft_p50_ms'],'ms/step',d['decode_ms_per_token_step'],'frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'],'cpu',d.get('cpu_baseline',{}).get('value'),d['clocks'], d['parity_check']['ok'], d['same_sample_e2e']['value'], 'launches', d['gpu_launches'])"
