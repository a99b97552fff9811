#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29714"
timeout 600 $TR bench.py --gpus 4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_tp4_bench.json 2> gpurun_out/r02_tp4_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02_tp4_bench.json'))
print('tp4', d['value'],'tok/s ttft',d['ttft_p50_ms'],'ms/step',d['decode_ms_per_token_step'],'prefill frac',d['roofline_prefill']['frac'], d['parity_check'])" || tail -5 gpurun_out/r02_tp4_bench.err
timeout 600 $TR bench.py --gpus 4 --model mixtral_8x7b --steps 3 --warmup 2 --no-cpu-baseline --no-parity-check > gpurun_out/r02_tp4_mixtral.json 2> gpurun_out/r02_tp4_mixtral.err
python -c "
import json;d=json.load(open('gpurun_out/r02_tp4_mixtral.json'))
print('mixtral tp4', d['value'],'tok/s ttft',d['ttft_p50_ms'],'ms/step',d['decode_ms_per_token_step'],'hbm frac',d['roofline']['frac'])" || tail -5 gpurun_out/r02_tp4_mixtral.err
