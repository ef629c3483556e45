mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -rs -k "tiny or host_class or colour or stride" ) > gpurun_out/pytest_new.log 2>&1
tail -15 gpurun_out/pytest_new.log
