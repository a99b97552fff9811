#!/bin/bash
MODE=continuous DURATION=${1:-30} timeout 600 python tools/batcher_load.py 2>&1 | grep -vE "^\s*$|Warning" | tail -5
