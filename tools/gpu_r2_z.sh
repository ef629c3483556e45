#!/bin/bash
# round 2, call z: ncu --set full of the level-0 launch of the volume refine kernel (mode 20) and of the default (mode 19), 4 pairs
mkdir -p gpurun_out
EPPM_REFINE_MODE=20 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_c2f_refine_vol -s 1 -c 1 -o gpurun_out/r2_refine_vol_l0 -f python tools/ncu_step.py 4 1 > gpurun_out/r2_ncu_refine_vol.log 2>&1; tail -2 gpurun_out/r2_ncu_refine_vol.log
EPPM_REFINE_MODE=19 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_c2f_refine_row -s 1 -c 1 -o gpurun_out/r2_refine_fastw_l0 -f python tools/ncu_step.py 4 1 > gpurun_out/r2_ncu_refine_fastw.log 2>&1; tail -2 gpurun_out/r2_ncu_refine_fastw.log
