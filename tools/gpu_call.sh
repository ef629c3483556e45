set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py 16 1 > gpurun_out/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_c2f_refine -c 2 -o gpurun_out/refine_tab -f python tools/ncu_step.py 4 1 > gpurun_out/ncu_refine.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_pm_search|k_prop_eval|k_prop_decide|k_flow_smooth' --launch-skip 40 -c 12 -o gpurun_out/pm -f python tools/ncu_step.py 4 1 > gpurun_out/ncu_pm.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json gpurun_out/bench_ref.json
