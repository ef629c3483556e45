#!/bin/bash
# BASELINE config 1: the shipped 640x480 pair through the UNMODIFIED main.cpp, three builds side by side:
#   oracle/_ref/runeppm            the reference (main.cpp + host class + its three .cu files)
#   oracle/_ref/runeppm_hostclass  the reference's main.cpp + host class on THIS library's stage functions (include/eppm_legacy_abi.h)
#   build/runeppm_b200             main.cpp on this library's class (include/compat)
# main.cpp prints the wall time of init + compute_flow ("GPU"): one cold call per process (context creation, module load, allocation).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
for exe in oracle/_ref/runeppm oracle/_ref/runeppm_hostclass build/runeppm_b200; do
  for k in 1 2 3; do
    d=$(mktemp -d); cp $ROOT/oracle/_ref/data/frame1*.ppm $d/
    ( cd $d && $ROOT/$exe 2>&1 | grep -i "GPU" | head -1 | sed "s|^|$exe run $k: |" )
    rm -rf $d
  done
done
