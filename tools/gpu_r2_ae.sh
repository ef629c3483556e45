#!/bin/bash
# round 2, call ae: refine inner loop at one sample per iteration: without the scheduling fence between the candidate blocks of the guard-free loop, with the tile row pointer hoisted
mkdir -p gpurun_out
run() { echo "$1"; env $2 timeout 150 python tools/variant_times.py 32 0 2>&1 | grep "^0 \|rror" | cut -c1-260; }
( for u in ju1 nofence hoist hoistnf; do run "$u" "EPPM_LIB_PATH=$PWD/build/ab/libeppm_b200_$u.so"; done ) | tee gpurun_out/r2_refine_loop_ab.txt
