#!/bin/bash
# Round-end check on one B200 (run as: gpurun --timeout 1500 -- 'bash tools/gpu_call.sh'): the GPU test suite, both bench arms, and the
# ncu launch list of a short bench run.  Everything lands in gpurun_out/; copy what should be judged into profiles/.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -rs ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json; cut -c1-200 gpurun_out/bench_ref.json
[ -n "$SKIP_LAUNCH_LIST" ] && exit 0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --batch 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
