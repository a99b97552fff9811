#!/bin/bash
timeout 400 python -m pytest tests/test_model_server_gpu.py tests/test_cb_gpu.py tests/test_sampling_gpu.py -m gpu -q -x --timeout 300 2>&1 | grep -E "passed|failed|^E  |Error" | head -12
