#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_model_server_gpu.py tests/test_tp_gpu.py -m gpu -q --timeout 500 2>&1 | grep -vE "^\s*$|Deprecation|importlib" | cut -c1-500 | tail -40
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_tp2.json 2> gpurun_out/bench_tp2.err
tail -5 gpurun_out/bench_tp2.err | cut -c1-400; cat gpurun_out/bench_tp2.json
