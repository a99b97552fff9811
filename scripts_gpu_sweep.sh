#!/bin/bash
N=${1:-2}
for v in "B200_PREFILL_GRID_MULT=2" "B200_PREFILL_GRID_MULT=4" "B200_PREFILL_GRID_MULT=8"; do
env $v timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/sw.json 2> gpurun_out/sw.err
python -c "
import json;d=json.load(open('gpurun_out/sw.json'))
print('tp$N $v', d['value'],'tok/s ttft',d['ttft_p50_ms'],'decode ms/step',d['decode_ms_per_token_step'])"
done
