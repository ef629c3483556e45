#!/bin/bash
# round 2, call w: refine mode 19 (fix-up-free loop behind an exact first patch row) after the vote-mask fix; mode 18 at 32 pairs: 15.367 ms per pair, hash e7ceccccf5cd9091
mkdir -p gpurun_out
for md in 19; do echo "refine mode $md"; EPPM_REFINE_MODE=$md timeout 150 python tools/variant_times.py 32 0 2>&1 | grep "^0 \|rror" | cut -c1-260; done | tee gpurun_out/r2_refine_fastw_ab2.txt
