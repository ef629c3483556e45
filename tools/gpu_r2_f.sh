#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rs -x ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log
for qt in 0 1 2; do
EPPM_SEARCH_QTEX=$qt timeout 600 python tools/variant_times.py 16 0 8192 > gpurun_out/r2_variant_times_f$qt.log 2>&1; echo qtex $qt; cut -c1-200 gpurun_out/r2_variant_times_f$qt.log
done
timeout 600 python tools/variant_times.py 16 524288 532480 > gpurun_out/r2_variant_times_f.log 2>&1; cut -c1-200 gpurun_out/r2_variant_times_f.log
