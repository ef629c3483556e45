#!/bin/bash
# round 2, final evidence: GPU suite, both bench arms, launch list of the bench command, ncu --set full of the level-0 refine launch
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -rs ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-700 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --batch 32 --distinct 4 --steps 1 --warmup 3 --no-cpu-baseline --no-reference-check > gpurun_out/launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_c2f_refine_row -s 1 -c 1 -o gpurun_out/r2_refine_final_l0 -f python tools/ncu_step.py 4 1 > gpurun_out/r2_ncu_refine_final.log 2>&1; tail -1 gpurun_out/r2_ncu_refine_final.log
