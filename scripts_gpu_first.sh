#!/bin/bash
# first GPU contact: kernel unit tests (one process per kernel family so a trap does not poison the rest),
# engine parity, smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for k in "gemm_store" "test_gemm_swiglu" "swapab_store" "swapab_swiglu" "splitk" "test_rmsnorm" "attn_prefill" "attn_decode" "rope_kv or argmax"; do
  echo "=== $k" >> gpurun_out/kernels.log
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "$k" 2>&1 | grep -vE "^\s*$" | tail -25 >> gpurun_out/kernels.log
done
grep -E "===|passed|failed|Error|error|assert" gpurun_out/kernels.log | cut -c1-300 | head -80
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 300 2>&1 | tail -80 > gpurun_out/engine.log
grep -E "passed|failed|Error|error|assert" gpurun_out/engine.log | cut -c1-300 | head -40
