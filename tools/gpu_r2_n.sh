#!/bin/bash
mkdir -p gpurun_out
run() { echo "rows=$1 cols=$2 variant=$3"; EPPM_SEARCH_ROWS=$1 EPPM_SEARCH_COLS=$2 timeout 600 python tools/variant_times.py 16 $3 2>&1 | cut -c1-120; }
run 8 0 0; run 16 0 0; run 16 16 0; run 8 32 0; run 8 0 1024; run 16 16 1024; run 8 0 512
