#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -rs -x ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu.log
for cfg in "0 0" "0 -5" "60 -5" "100 -5" "60 0" "36 -5"; do
timeout 300 python tools/overlap_probe.py 16 $cfg 2>&1 | tail -1
done | tee gpurun_out/r2_overlap_probe.log
