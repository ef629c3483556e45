mkdir -p gpurun_out
SS_NO_REF=1 SS_STRIDES=1,3 timeout 600 python tools/stream_sweep.py 300 60 > gpurun_out/stream_sweep13.log 2>&1
tail -19 gpurun_out/stream_sweep13.log | cut -c1-200
