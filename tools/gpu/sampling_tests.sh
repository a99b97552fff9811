#!/bin/bash
timeout 600 python -m pytest tests/test_sampling_gpu.py -m gpu -q -x --timeout 400 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | tail -25
