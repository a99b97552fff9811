#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 120 -k "rope_kv" 2>&1 | grep -E "AssertionError|passed|failed" | cut -c1-600 | head -5
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q --timeout 500 -k "llama3_8b" 2>&1 | tail -15 | cut -c1-400
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -5 gpurun_out/bench1.err; cat gpurun_out/bench1.json
cat gpurun_out/parity_stats.json | tr -d '\n' | cut -c1-1500
