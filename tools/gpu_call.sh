mkdir -p gpurun_out
( timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "fixture or tiny or uncalled or upsampling or image_smoothing or colour" ) > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer_memcheck.log | tail -8
( timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "subpixel_refine_against_committed" ) > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitizer_racecheck.log | tail -6
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_pm_search_joint|k_flow_smooth4' --launch-skip 2 -c 3 -o gpurun_out/search_smooth -f python tools/ncu_step.py 4 1 > gpurun_out/ncu_ss.log 2>&1
tail -2 gpurun_out/ncu_ss.log
