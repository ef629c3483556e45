set -x
mkdir -p gpurun_out
timeout 300 python tools/gen_golden_subpix.py > gpurun_out/gen_subpix.log 2>&1
tail -2 gpurun_out/gen_subpix.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --batch 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/launches_bench.log | cut -c1-300
