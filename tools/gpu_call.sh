set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
head -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/stream_sweep.py 300 60 > gpurun_out/stream_sweep.log 2>&1
tail -32 gpurun_out/stream_sweep.log | cut -c1-330
