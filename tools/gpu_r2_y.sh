#!/bin/bash
# round 2, call y: refine with the shared AD + census volume (EPPM_REFINE_MODE 20 / 21 / 22 = 6 / 7 / 5 CTAs per SM) against mode 19 (15.09 ms per pair, hash e7ceccccf5cd9091 at 32 pairs)
mkdir -p gpurun_out
for md in ${MODES:-20 21 22}; do echo "refine mode $md"; EPPM_REFINE_MODE=$md timeout 120 python tools/variant_times.py 32 0 2>&1 | grep "^0 \|rror\|Traceback" | cut -c1-260; done | tee gpurun_out/r2_refine_vol_ab.txt
