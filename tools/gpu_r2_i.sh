#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -rs -x ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu.log
for md in 0 8 9 10; do
EPPM_REFINE_MODE=$md timeout 600 python tools/variant_times.py 16 0 > gpurun_out/r2_variant_times_i$md.log 2>&1; echo refine mode $md; cut -c1-200 gpurun_out/r2_variant_times_i$md.log
done
