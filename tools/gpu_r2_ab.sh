#!/bin/bash
# round 2, call ab: small knobs at 32 pairs per launch -- refine mode 23 (guard-free fix-up-free loop), propagation queue items per warp / L1 carve-out
mkdir -p gpurun_out
run() { echo "$1"; env $2 timeout 150 python tools/variant_times.py 32 0 2>&1 | grep "^0 \|rror" | cut -c1-260; }
( run "default" "X=1"; run "refine mode 23" "EPPM_REFINE_MODE=23"; run "prop batch 16" "EPPM_PROP_BATCH=16"; run "prop carveout 25" "EPPM_PROP_CARVEOUT=25"; run "prop carveout 60" "EPPM_PROP_CARVEOUT=60" ) | tee gpurun_out/r2_knobs32.txt
