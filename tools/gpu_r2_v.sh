#!/bin/bash
# round 2, call v: refine without the __expf fix-up test behind an exact first patch row (EPPM_REFINE_MODE=19) against the default (18), 32-pair chunks;
# baoCudaPatchMatch_Scaled against the reference build; the whole GPU suite with mode 19 (its refine tests compare with the reference bit for bit)
mkdir -p gpurun_out
for md in 18 19; do echo "refine mode $md"; EPPM_REFINE_MODE=$md timeout 600 python tools/variant_times.py 32 0 2>&1 | grep "^0 " | cut -c1-260; done | tee gpurun_out/r2_refine_fastw_ab.txt
timeout 300 python -m pytest tests -m gpu -q -x -k "scaled" 2>&1 | tail -15
( time EPPM_REFINE_MODE=19 timeout 900 python -m pytest tests -m gpu -q -rs ) > gpurun_out/pytest_gpu_mode19.log 2>&1; tail -8 gpurun_out/pytest_gpu_mode19.log
