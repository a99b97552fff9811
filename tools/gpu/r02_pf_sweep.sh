#!/bin/bash
# weight-stream prefetcher: correctness subset + A/B + lead sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_cb_gpu.py -m gpu -q --timeout 600 -x -k "not headline" > gpurun_out/r02_pf_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02_pf_tests.log | cut -c1-300
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity-check"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 $B > gpurun_out/r02_pf_$name.json 2> gpurun_out/r02_pf_$name.err
  python -c "
import json;d=json.load(open('gpurun_out/r02_pf_$name.json'))
print('$name', d['value'],'tok/s ttft',d['ttft_p50_ms'],'ms/step',d['decode_ms_per_token_step'],'frac',d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -5 gpurun_out/r02_pf_$name.err
}
run off B200_NO_PREFETCHER=1
run lead12 B200_PF_LEAD=12
run lead6 B200_PF_LEAD=6
run lead24 B200_PF_LEAD=24
run lead48 B200_PF_LEAD=48
run lead96 B200_PF_LEAD=96
