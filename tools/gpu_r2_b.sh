#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rs -x ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log
# 0 = new default (warp-cooperative prop eval + search); 8192 thread prop; 32768 thread search; 40960 = round-1 default
timeout 600 python tools/variant_times.py 8 0 8192 32768 40960 > gpurun_out/r2_variant_times_b.log 2>&1; cut -c1-300 gpurun_out/r2_variant_times_b.log
