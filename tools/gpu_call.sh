set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -s -k "uncalled or subpixel or variant_switches" ) > gpurun_out/pytest_new.log 2>&1
tail -30 gpurun_out/pytest_new.log
