#!/bin/bash
# round 2, call t: refine with the image-1 tile staged by TMA (EPPM_REFINE_MODE=18) against the default (10); search instantiations for the 256-thread tiles
mkdir -p gpurun_out
for md in 10 18; do echo "refine mode $md"; EPPM_REFINE_MODE=$md timeout 600 python tools/variant_times.py 16 0 2>&1 | grep "^0 " | cut -c1-260; done
for sm in 0 1 2 3 4; do echo "search 256 mode $sm"; EPPM_SEARCH_256MODE=$sm timeout 600 python tools/variant_times.py 16 0 2>&1 | grep "^0 " | cut -c1-260; done
