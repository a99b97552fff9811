#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_gpu.py -m gpu -q --timeout 800 -k "tp4_tp8" > gpurun_out/r02_tp48_tests_final.log 2>&1
echo "tp4/8 tests rc=$?"; tail -3 gpurun_out/r02_tp48_tests_final.log | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29716"
timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_tp8_final.json 2> gpurun_out/r02_tp8_final.err
python -c "
import json;d=json.load(open('gpurun_out/r02_tp8_final.json'))
print('tp8', d['value'],'tok/s e2e', d['e2e']['value'], 'ttft',d['ttft_p50_ms'],'ms/step',d['decode_ms_per_token_step'],'prefill frac',d['roofline_prefill']['frac'], d['parity_check'])" || tail -5 gpurun_out/r02_tp8_final.err
