#!/bin/bash
# ncu evidence for round 2: launch list of one full request (prefill + 2 decode steps) + full-set captures per kernel family
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --gen-len 3 --no-cpu-baseline --no-parity-check"
ALL='regex:gemm_tn|attn_|rmsnorm|rope_kv|argmax|step_update|embed_gather|pack_|gather_rows|swiglu_reduce'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$ALL" -c 1400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/pf0.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r02_launches.csv)"
cap() { # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o gpurun_out/r02_$1 $B > gpurun_out/pf_$1.log 2>&1
  echo "$1 rc=$?"
}
cap gemm2cta gemm_tn_2cta 64 4
cap gemm_decode 'gemm_tn_kernel' 1 5
cap attn 'attn_prefill_tc|attn_decode' 31 2
cap small 'rmsnorm_rows|rope_kv|rmsnorm_kernel' 66 4
ls -la gpurun_out/ | grep "r02_.*ncu-rep"
