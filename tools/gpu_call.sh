set -x
mkdir -p gpurun_out
timeout 300 python tools/variant_times.py 16 0 512 1024 65536 > gpurun_out/variant_times.log 2>&1
for n in nopopc noex2 noboth; do
  EPPM_LIB_PATH=$PWD/build/whatif/libeppm_b200_$n.so timeout 300 python tools/variant_times.py 16 0 > gpurun_out/whatif_$n.log 2>&1
done
cat gpurun_out/variant_times.log gpurun_out/whatif_*.log
