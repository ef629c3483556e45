#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rs -x -k "variant_switches or every_pass or stride" ) > gpurun_out/r2_pytest_gpu_c.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu_c.log
for cv in 40 20 70; do
EPPM_PROP_CARVEOUT=$cv timeout 600 python tools/variant_times.py 16 0 > gpurun_out/r2_variant_times_c$cv.log 2>&1; echo carve $cv; cut -c1-200 gpurun_out/r2_variant_times_c$cv.log
done
timeout 600 python tools/variant_times.py 16 8192 131072 > gpurun_out/r2_variant_times_c.log 2>&1; cut -c1-200 gpurun_out/r2_variant_times_c.log
