#!/bin/bash
# round 2, call x: the whole GPU suite on the round's final kernels (refine mode 19, 32 pairs per launch, baoCudaPatchMatch_Scaled), both bench arms, launch list
mkdir -p gpurun_out
( time timeout 840 python -m pytest tests -m gpu -q -rs -x ) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-700 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --batch 32 --distinct 4 --steps 1 --warmup 3 --no-cpu-baseline --no-reference-check > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-300
