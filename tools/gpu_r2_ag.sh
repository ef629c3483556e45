#!/bin/bash
# round 2, call ag: unroll factors of the smoothing kernel's tap loop (3 = default) and of the random search's sample loop (2 = default)
mkdir -p gpurun_out
run() { echo "$1"; env $2 timeout 150 python tools/variant_times.py 32 0 2>&1 | grep "^0 \|rror" | cut -c1-260; }
( for u in base s4u1 s4u7 sj1 sj5; do run "$u" "EPPM_LIB_PATH=$PWD/build/ab/libeppm_b200_$u.so"; done ) | tee gpurun_out/r2_unroll_ab.txt
