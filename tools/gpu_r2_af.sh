#!/bin/bash
# round 2, call af: refine row loop -- __syncwarp as the scheduling fence between candidate blocks, guard-free exact first row
mkdir -p gpurun_out
run() { echo "$1"; env $2 timeout 150 python tools/variant_times.py 32 0 2>&1 | grep "^0 \|rror" | cut -c1-260; }
( for u in base syncwarp firstrow; do run "$u" "EPPM_LIB_PATH=$PWD/build/ab/libeppm_b200_$u.so"; done ) | tee gpurun_out/r2_refine_loop_ab2.txt
