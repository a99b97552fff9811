#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cb_gpu.py tests/test_sampling_gpu.py tests/test_model_server_gpu.py -m gpu -q --timeout 500 2>&1 | tail -4 | cut -c1-300
MODE=continuous SLOTS=512 DURATION=40 timeout 600 python tools/http_load.py 2> gpurun_out/r02_http_load_continuous.err | tail -1 | cut -c1-1200
tail -3 gpurun_out/r02_http_load_continuous.err | cut -c1-300
MODE=continuous SLOTS=64 DURATION=30 timeout 600 python tools/http_load.py 2> gpurun_out/r02_http_load_continuous64.err | tail -1 | cut -c1-1200
cp gpurun_out/r02_http_load_continuous.json gpurun_out/r02_http_load_continuous64.json 2>/dev/null
