#!/bin/bash
# ncu evidence on the final round-2 code: launch list of one request + one full-set pass over the first prefill layer
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --gen-len 3 --no-cpu-baseline --no-parity-check"
ALL='regex:gemm_tn|attn_|rmsnorm|rope_kv|argmax|step_update|embed_gather|pack_|gather_rows|swiglu_reduce'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "$ALL" -c 1400 --csv --log-file gpurun_out/r02_final_launches.csv $B > gpurun_out/pf0.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r02_final_launches.csv)"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:gemm_tn_2cta|attn_prefill_tc" -s 5 -c 5 -o gpurun_out/r02_final_layer $B > gpurun_out/pf_layer.log 2>&1
echo "layer rc=$?"
ls -la gpurun_out/ | grep "r02_final"
