#!/bin/bash
mkdir -p gpurun_out
for rows in 4 1 2 8; do
EPPM_SEARCH_ROWS=$rows timeout 600 python tools/variant_times.py 16 0 > gpurun_out/r2_variant_times_m$rows.log 2>&1; echo search rows $rows; cut -c1-200 gpurun_out/r2_variant_times_m$rows.log
done
timeout 600 python -m pytest tests -m gpu -q -x -k "every_pass or variant_switches or tiny_and_extreme or patch_stride or non_default_depth" 2>&1 | tail -3
