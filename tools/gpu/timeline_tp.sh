#!/bin/bash
N=${1:-4}
env MODEL=${2:-llama3_8b} LAYERS=4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tools/timeline.py 2>&1 | grep -vE "^\*|OMP_NUM|^\s*$|Warning" | sed -n '2,3p;14,44p'
