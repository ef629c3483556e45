#!/bin/bash
# compute-sanitizer over the GPU test suite (run under gpurun).  memcheck: every test except the 1080p ones (minutes each under the tool);
# racecheck: this library's kernels only (kernel names k_*; the reference's d_update_random_guess has shared-memory hazards of its own that
# would drown the report), on the tests that drive the propagation queue, the weighted median, the refine exchange (default kernel, its
# fix-up-free loop on adversarial inputs, and the shared-volume variant with its per-row barriers), the scaled PatchMatch and the TMA smoothing.
mkdir -p gpurun_out
export EPPM_UNDER_SANITIZER=1
( time timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -q -k "not full_hd" ) > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -6 gpurun_out/r2_sanitizer_memcheck.log
( time timeout 900 compute-sanitizer --tool racecheck --kernel-name kns=k_ --error-exitcode 3 python -m pytest tests -m gpu -q -k "every_pass or consistency_and_c2f or gpu_vs_cpu_oracle_small or batch_equals_single or video_stream_reuses or adversarial or patchmatch_scaled" ) > gpurun_out/r2_sanitizer_racecheck.log 2>&1; tail -6 gpurun_out/r2_sanitizer_racecheck.log
( time EPPM_VARIANT=8388608 timeout 600 compute-sanitizer --tool racecheck --kernel-name kns=k_c2f --error-exitcode 3 python -m pytest tests -m gpu -q -k "gpu_vs_cpu_oracle_small or batch_equals_single" ) > gpurun_out/r2_sanitizer_racecheck_vol.log 2>&1; tail -6 gpurun_out/r2_sanitizer_racecheck_vol.log
