#!/bin/bash
# round 2, call u: pairs per launch (context max_batch) swept: the propagation's 800 dependent launches per chunk are latency bound
mkdir -p gpurun_out
for n in 16 32 48 64; do echo "chunk $n"; timeout 600 python tools/variant_times.py $n 0 2>&1 | grep "^0 " | cut -c1-260; done | tee gpurun_out/r2_chunk_sweep.txt
timeout 400 python bench.py --chunk 32 --steps 2 --warmup 3 --no-cpu-baseline --no-reference-check > gpurun_out/r2_bench_chunk32.json 2> gpurun_out/r2_bench_chunk32.err; cut -c1-900 gpurun_out/r2_bench_chunk32.json
timeout 400 python bench.py --chunk 64 --steps 2 --warmup 3 --no-cpu-baseline --no-reference-check > gpurun_out/r2_bench_chunk64.json 2> gpurun_out/r2_bench_chunk64.err; cut -c1-900 gpurun_out/r2_bench_chunk64.json
