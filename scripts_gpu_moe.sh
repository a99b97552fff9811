#!/bin/bash
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -x --timeout 600 -k "moe" 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | tail -4
MODEL=mixtral_8x7b LAYERS=3 timeout 600 python tools/timeline.py 2>&1 | grep -vE "^\*|OMP_NUM|^\s*$|Warning" | sed -n '1,2p;16,30p'
