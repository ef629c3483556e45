#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prop_eval_r -s 230 -c 3 -o gpurun_out/r2_prop_eval_r -f python tools/ncu_step.py 16 1 > gpurun_out/r2_ncu_prop_r.log 2>&1; tail -3 gpurun_out/r2_ncu_prop_r.log
EPPM_VARIANT=8192 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prop_eval -s 230 -c 3 -o gpurun_out/r2_prop_eval_t -f python tools/ncu_step.py 16 1 > gpurun_out/r2_ncu_prop_t.log 2>&1; tail -3 gpurun_out/r2_ncu_prop_t.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pm_search_joint -s 4 -c 1 -o gpurun_out/r2_search_t -f python tools/ncu_step.py 16 1 > gpurun_out/r2_ncu_search_t.log 2>&1; tail -3 gpurun_out/r2_ncu_search_t.log
