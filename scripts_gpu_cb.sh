#!/bin/bash
timeout 600 python -m pytest tests/test_cb_gpu.py -m gpu -q -x --timeout 300 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | tail -25
