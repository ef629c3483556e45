#!/bin/bash
# Timing-experiment builds of the library (WRONG results by construction; never loaded by tests or bench):
#   tools/build_whatif.sh NAME -DEPPM_WHATIF_...   ->  build/whatif/libeppm_b200_NAME.so
set -e
cd "$(dirname "$0")/../eppm_b200/csrc"
name=$1; shift
out=../../build/whatif; mkdir -p $out/obj_$name
for f in context prepare patchmatch consistency refine legacy_abi bao_class; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include "$@" -c $f.cu -o $out/obj_$name/$f.o 2>/dev/null &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xlinker -Bsymbolic -o $out/libeppm_b200_$name.so $out/obj_$name/*.o 2>/dev/null
echo built $out/libeppm_b200_$name.so
