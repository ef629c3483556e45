#!/bin/bash
mkdir -p gpurun_out
for md in 8 9 10 11; do
EPPM_REFINE_MODE=$md timeout 600 python tools/variant_times.py 16 0 > gpurun_out/r2_variant_times_j$md.log 2>&1; echo refine mode $md; cut -c1-200 gpurun_out/r2_variant_times_j$md.log
done
