#!/bin/bash
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_kernels_gpu.py -m gpu -q -x --timeout 600 -k "moe or gemm or Gemm" 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | tail -8
