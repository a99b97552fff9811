#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/final_tests.log 2>&1
echo "tests rc=$?"; grep -vE "^\s*$|Deprecation|importlib|warnings" gpurun_out/final_tests.log | tail -6 | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_final.json'))
print('N=1', d['value'],'tok/s e2e',d['e2e']['value'],'ttft',d['ttft_p50_ms'],'ms/step',d['decode_ms_per_token_step'],'frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'],'cpu',d.get('cpu_baseline'),d['clocks'])"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
B="python bench.py --steps 1 --warmup 1 --gen-len 3 --no-cpu-baseline"
ALL='regex:gemm_tn|attn_|rmsnorm|rope_kv|argmax|step_update|embed_gather|pack_|gather_rows|swiglu_reduce'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$ALL" -c 1200 --csv --log-file gpurun_out/r01g_launches.csv $B > gpurun_out/pf0.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r01g_launches.csv)"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:attn_prefill_tc" -s 31 -c 1 -o gpurun_out/r01g_attn_tc $B > gpurun_out/pf_attn.log 2>&1
echo "attn rc=$?"
ls -la gpurun_out | grep r01g
