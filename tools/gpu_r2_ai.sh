#!/bin/bash
# round 2, call ai: default refine kernel compiled for 8 / 7 (default) / 6 CTAs per SM
mkdir -p gpurun_out
run() { echo "$1"; env $2 timeout 150 python tools/variant_times.py 32 0 2>&1 | grep "^0 \|rror" | cut -c1-260; }
( run "7 CTAs (default)" "X=1"; run "8 CTAs" "EPPM_REFINE_MODE=24"; run "6 CTAs" "EPPM_REFINE_MODE=25" ) | tee gpurun_out/r2_refine_minb_ab.txt
