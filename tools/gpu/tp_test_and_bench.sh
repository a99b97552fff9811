#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
if [ "$2" = "test" ]; then
timeout 600 python -m pytest tests/test_tp_gpu.py -m gpu -q --timeout 500 > gpurun_out/tp_test.log 2>&1
grep -vE "^\s*$|Deprecation|importlib" gpurun_out/tp_test.log | cut -c1-900 | tail -12
fi
for v in $3 ""; do
env $v timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_tp${N}.json 2> gpurun_out/bench_tp${N}.err
grep -iE "error|Traceback|timeout" gpurun_out/bench_tp${N}.err | head -5
python -c "
import json;d=json.load(open('gpurun_out/bench_tp${N}.json'))
print('tp$N $v', d['value'],'tok/s ttft',d['ttft_p50_ms'],'decode ms/step',d['decode_ms_per_token_step'],'hbm frac',d['roofline']['frac'],'prefill frac',d['roofline_prefill']['frac'])"
done
