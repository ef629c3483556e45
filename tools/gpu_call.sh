mkdir -p gpurun_out
timeout 300 python tools/tiny_dev.py > gpurun_out/tiny.log 2>&1; cat gpurun_out/tiny.log | tail -30
