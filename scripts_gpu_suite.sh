#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -40 > gpurun_out/suite.log
grep -vE "^\s*$" gpurun_out/suite.log | cut -c1-400 | tail -40
