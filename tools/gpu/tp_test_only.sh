#!/bin/bash
timeout 500 python -m pytest tests/test_tp_gpu.py -m gpu -q --timeout 400 2>&1 | grep -E "passed|failed|Error|assert" | tail -5 | cut -c1-300
