#!/bin/bash
# ncu evidence: launch list of one full step + full-set captures of the dominant kernels
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --gen-len 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/b_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches.csv)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 129 -c 5 -o gpurun_out/prof_gemm_decode $B > gpurun_out/p1.log 2>&1
echo "gemm decode rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 0 -c 4 -o gpurun_out/prof_gemm_prefill $B > gpurun_out/p2.log 2>&1
echo "gemm prefill rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 31 -c 3 -o gpurun_out/prof_attn $B > gpurun_out/p3.log 2>&1
echo "attn rc=$?"
ls -la gpurun_out/
