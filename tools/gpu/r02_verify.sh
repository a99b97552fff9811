#!/bin/bash
# round 2, first GPU pass: all -m gpu tests, a short bench of both arms, timeline of one decode layer
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_tests.log 2>&1
echo "tests rc=$?"; grep -vE "^\s*$|Deprecation|importlib|warnings" gpurun_out/r02_tests.log | tail -8 | cut -c1-400
cp gpurun_out/parity_stats.json gpurun_out/r02_parity_stats.json 2>/dev/null
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r02_bench.json'))
print('N=1', d['value'],'tok/s e2e',d['e2e']['value'],'ttft',d['ttft_p50_ms'],'ms/step',d['decode_ms_per_token_step'],'frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'],'cpu',d.get('cpu_baseline'),d['clocks'], d['parity_check'], d['same_sample_e2e'])"
tail -3 gpurun_out/r02_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-900 gpurun_out/r02_bench_ref.json
