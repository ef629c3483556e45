#!/bin/bash
mkdir -p gpurun_out
for md in 10 16 17; do echo "refine mode $md"; EPPM_REFINE_MODE=$md timeout 600 python tools/variant_times.py 16 0 2>&1 | cut -c1-130; done
