#!/bin/bash
N=${1:-8}
for v in "" "B200_GU_STREAMK=1"; do
echo "== $v"
env $v LAYERS=4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tools/timeline.py 2>&1 | grep -vE "^\*|OMP_NUM|^\s*$|Warning" | sed -n '2p;22,32p'
done
