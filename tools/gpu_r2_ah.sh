#!/bin/bash
# round 2, call ah: scheduling fence of the guard-free refine loop in front of candidate blocks 0..2 (default), 1..2, 2 only
mkdir -p gpurun_out
run() { echo "$1"; env $2 timeout 150 python tools/variant_times.py 32 0 2>&1 | grep "^0 \|rror" | cut -c1-260; }
( for u in base ff1 ff2; do run "$u" "EPPM_LIB_PATH=$PWD/build/ab/libeppm_b200_$u.so"; done ) | tee gpurun_out/r2_fence_ab.txt
