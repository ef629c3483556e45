#!/bin/bash
# round 2, first GPU call: does the new default (packed refine, chain propagation) pass parity, and what does it buy?
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rs -x ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu.log
timeout 120 build/peak_alu > gpurun_out/r2_peak_alu.log 2>&1; tail -12 gpurun_out/r2_peak_alu.log
# variants: 0 = new default; 2048 = scalar refine; 8192 = queue propagation; 10240 = round-1 default; 4096 = packed refine w/ branch
timeout 600 python tools/variant_times.py 8 0 2048 8192 10240 4096 > gpurun_out/r2_variant_times.log 2>&1; cat gpurun_out/r2_variant_times.log | cut -c1-300
timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_first.json 2> gpurun_out/r2_bench_first.err; cut -c1-600 gpurun_out/r2_bench_first.json
