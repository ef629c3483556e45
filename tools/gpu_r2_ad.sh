#!/bin/bash
# round 2, call ad: samples per iteration of the refine's inner loop (RF_JUNROLL = 1 / 2 (default) / 5 / 10), separate builds under build/ab
mkdir -p gpurun_out
run() { echo "$1"; env $2 timeout 150 python tools/variant_times.py 32 0 2>&1 | grep "^0 \|rror" | cut -c1-260; }
( run "junroll 2 (default build)" "X=1"; for u in 1 5 10; do run "junroll $u" "EPPM_LIB_PATH=$PWD/build/ab/libeppm_b200_ju$u.so"; done ) | tee gpurun_out/r2_junroll.txt
