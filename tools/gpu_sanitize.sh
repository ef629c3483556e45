#!/bin/bash
# compute-sanitizer over the GPU test suite (run under gpurun).  memcheck: every test except the 1080p ones (minutes each under the tool);
# racecheck: this library's kernels only (kernel names k_*; the reference's d_update_random_guess has shared-memory hazards of its own that
# would drown the report), on the tests that drive the propagation queue, the weighted median, the refine exchange and the TMA smoothing.
mkdir -p gpurun_out
export EPPM_UNDER_SANITIZER=1
( time timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -q -k "not full_hd" ) > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -6 gpurun_out/r2_sanitizer_memcheck.log
( time timeout 1500 compute-sanitizer --tool racecheck --kernel-name kns=k_ --error-exitcode 3 python -m pytest tests -m gpu -q -k "every_pass or consistency_and_c2f or gpu_vs_cpu_oracle_small or batch_equals_single or video_stream_reuses" ) > gpurun_out/r2_sanitizer_racecheck.log 2>&1; tail -6 gpurun_out/r2_sanitizer_racecheck.log
