#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "every_pass or variant_switches or tiny_and_extreme or patch_stride or non_default_depth or video_stream or full_hd_vs" 2>&1 | tail -3
echo "batch 16"; timeout 600 python tools/variant_times.py 16 0 2097152 8192 2>&1 | cut -c1-120
echo "batch 8"; EPPM_PROP_BATCH=8 timeout 600 python tools/variant_times.py 16 0 2097152 2>&1 | cut -c1-120
echo "32 pairs"; timeout 600 python tools/variant_times.py 32 0 2>&1 | cut -c1-120
