#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -rs -x ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2_pytest_gpu.log
for md in 0 1 2 3 4 5 6 7; do
EPPM_REFINE_MODE=$md timeout 600 python tools/variant_times.py 16 0 > gpurun_out/r2_variant_times_g$md.log 2>&1; echo refine mode $md; cut -c1-200 gpurun_out/r2_variant_times_g$md.log
done
timeout 900 python tools/parity_e2e.py > gpurun_out/r2_parity_e2e.log 2>&1; cut -c1-900 gpurun_out/r2_parity_e2e.log
timeout 300 python tools/parity_stages.py x > gpurun_out/r2_parity_stages_snapshot.log 2>&1; cp gpurun_out/parity_stages.json gpurun_out/r2_parity_stages_snapshot.json
EPPM_INPLACE_LEGACY=1 timeout 300 python tools/parity_stages.py x > gpurun_out/r2_parity_stages_inplace.log 2>&1; cp gpurun_out/parity_stages.json gpurun_out/r2_parity_stages_inplace.json
grep -E "Outlier|WMF iters=20|FlowSmoothing" gpurun_out/r2_parity_stages_snapshot.log | cut -c1-200
echo INPLACE; grep -E "Outlier|WMF iters=20|FlowSmoothing" gpurun_out/r2_parity_stages_inplace.log | cut -c1-200
