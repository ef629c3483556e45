#!/bin/bash
# round 2, call ac: refine default = mode 23 + per-sample spatial weight table; GPU suite
mkdir -p gpurun_out
run() { echo "$1"; env $2 timeout 150 python tools/variant_times.py 32 0 2>&1 | grep "^0 \|rror" | cut -c1-260; }
( run "default + weighted median with three candidates per window pass" "X=1" ) | tee gpurun_out/r2_knobs32b.txt
( time timeout 600 python -m pytest tests -m gpu -q -rs -x ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
