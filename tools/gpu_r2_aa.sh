#!/bin/bash
# round 2, call aa: census + pack with a TMA-staged tile (default) against the per-thread loads (EPPM_VARIANT=16777216): prepare-stage time, same flow bits;
# GPU suite on the result; colour wheel exactness probe
mkdir -p gpurun_out
timeout 200 python tools/variant_times.py 32 0 16777216 2>&1 | grep "^0 \|^16777216 \|rror" | cut -c1-260 | tee gpurun_out/r2_census_tma_ab.txt
( time timeout 600 python -m pytest tests -m gpu -q -rs -x ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
EPPM_TEST_COLOUR_EXACT=1 timeout 100 python -m pytest tests -m gpu -q -s -k colour 2>&1 | grep "colour coding\|passed\|failed" | head
