#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
timeout 400 $TR tools/tiled_run_c.py > gpurun_out/r2_tiled_c_${N}gpu.log 2>&1; grep "^{" gpurun_out/r2_tiled_c_${N}gpu.log | cut -c1-400
timeout 400 $TR bench.py --config 4 --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_cfg4_${N}gpu.json 2> gpurun_out/r2_bench_cfg4_${N}gpu.err; grep "^{" gpurun_out/r2_bench_cfg4_${N}gpu.json | cut -c1-700; tail -2 gpurun_out/r2_bench_cfg4_${N}gpu.err
