#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --model mixtral_8x7b --steps 2 --warmup 3 > gpurun_out/bench_mixtral_tp${N}.json 2> gpurun_out/bench_mixtral_tp${N}.err
echo "rc=$?"
grep -iE "error|Traceback|timeout|b200:" gpurun_out/bench_mixtral_tp${N}.err | head -8
tail -3 gpurun_out/bench_mixtral_tp${N}.err | cut -c1-300
python -c "
import json;d=json.load(open('gpurun_out/bench_mixtral_tp${N}.json'))
print('mixtral tp$N', d['value'],'tok/s ttft',d['ttft_p50_ms'],'decode ms/step',d['decode_ms_per_token_step'],'hbm frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'], 'e2e', d['e2e']['value'])"
