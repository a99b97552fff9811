#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29715"
for i in 1 2; do
for cfg in "fold X=1" "combinek B200_ATTN_COMBINE_KERNEL=1"; do
  set -- $cfg; name=$1; shift
  env "$@" LAYERS=8 ROWS=4 timeout 300 $TR tools/timeline.py 2>&1 | grep "step span" | sed "s/^/$name: /"
done
done
