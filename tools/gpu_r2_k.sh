#!/bin/bash
# measurement evidence of round 2: bench (both arms), launch list of the bench command, full ncu capture of the dominant kernel, sanitizers
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cut -c1-300 gpurun_out/r2_bench_ref.json
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; cut -c1-2500 gpurun_out/r2_bench.json; tail -2 gpurun_out/r2_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --batch 16 --distinct 4 --steps 1 --warmup 3 --no-cpu-baseline --no-reference-check > gpurun_out/r2_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_c2f_refine_row -s 1 -c 1 -o gpurun_out/r2_refine_row_l0 -f python tools/ncu_step.py 4 1 > gpurun_out/r2_ncu_refine_row.log 2>&1; tail -2 gpurun_out/r2_ncu_refine_row.log
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -q -x -k "not full_hd and not variant_switches" ) > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -8 gpurun_out/r2_sanitizer_memcheck.log
( time timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests -m gpu -q -x -k "every_pass or consistency_and_c2f or gpu_vs_cpu_oracle_small or batch_equals_single" ) > gpurun_out/r2_sanitizer_racecheck.log 2>&1; tail -8 gpurun_out/r2_sanitizer_racecheck.log
