#!/bin/bash
lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-200 | head -3
lscpu | grep -oE "amx_bf16|avx512_bf16|avx512f" | sort -u | tr '\n' ' '; echo
python -c "
import sys; sys.path.insert(0,'.')
import bench, os
print('cpu_count', os.cpu_count(), 'usable', bench.usable_cores(), 'picked', bench.pick_cpu_threads())"
cat /sys/fs/cgroup/cpu.max 2>/dev/null
time (timeout 400 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | cut -c1-700)
