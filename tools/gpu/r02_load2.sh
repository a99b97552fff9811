#!/bin/bash
mkdir -p gpurun_out
MODE=continuous SLOTS=512 KV_PAGES=2000 DURATION=30 timeout 600 python tools/http_load.py 2> gpurun_out/r02_http_load_kvtier.err | tail -1 | cut -c1-1400
tail -2 gpurun_out/r02_http_load_kvtier.err | cut -c1-300
MODE=continuous SLOTS=512 CHUNK=8192 DURATION=30 timeout 600 python tools/http_load.py 2> gpurun_out/r02_http_load_c8192.err | tail -1 | cut -c1-1400
bash tools/gpu/r02_verify.sh
