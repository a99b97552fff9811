#!/bin/bash
./scripts_gpu_cfg4.sh 4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_tp4.json 2> gpurun_out/bench_tp4.err
python -c "
import json;d=json.load(open('gpurun_out/bench_tp4.json'))
print('llama tp4', d['value'],'tok/s ttft',d['ttft_p50_ms'],'decode ms/step',d['decode_ms_per_token_step'],'hbm frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'])"
