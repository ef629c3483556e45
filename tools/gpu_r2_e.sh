#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rs -x ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log
# 0 = coop round-major + memo; 8192 = thread eval + memo; 270336 = thread eval, no memo (round 1)
timeout 600 python tools/variant_times.py 16 0 8192 270336 262144 > gpurun_out/r2_variant_times_e.log 2>&1; cut -c1-200 gpurun_out/r2_variant_times_e.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cut -c1-1500 gpurun_out/r2_bench_ref.json; tail -3 gpurun_out/r2_bench_ref.err
REF_SHIM_NOCACHE=1 timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_nocache.json 2> gpurun_out/r2_bench_ref_nocache.err; cut -c1-300 gpurun_out/r2_bench_ref_nocache.json
