#!/bin/bash
# 2-GPU checks: library-side tiling (bit-identity + time), strong-scaling bench of both arms, config 4 bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR tools/tiled_run_c.py > gpurun_out/r2_tiled_c_2gpu.log 2>&1; tail -3 gpurun_out/r2_tiled_c_2gpu.log | cut -c1-400
timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; cut -c1-700 gpurun_out/r2_bench_2gpu.json; tail -2 gpurun_out/r2_bench_2gpu.err
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_2gpu.json 2> gpurun_out/r2_bench_ref_2gpu.err; cut -c1-300 gpurun_out/r2_bench_ref_2gpu.json; tail -2 gpurun_out/r2_bench_ref_2gpu.err
timeout 600 $TR bench.py --config 4 --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_cfg4_2gpu.json 2> gpurun_out/r2_bench_cfg4_2gpu.err; cut -c1-900 gpurun_out/r2_bench_cfg4_2gpu.json; tail -2 gpurun_out/r2_bench_cfg4_2gpu.err
