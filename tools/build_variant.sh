#!/bin/bash
# Build a tuning variant of libeppm_b200.so into build/ab/libeppm_b200_<name>.so with extra nvcc flags (e.g. -DRF_MINBLOCKS=7).
# usage: tools/build_variant.sh <name> "<extra nvcc flags>"
set -e
name=$1; extra=$2
root=$(cd "$(dirname "$0")/.." && pwd)
obj=$root/build/ab/obj_$name
mkdir -p $obj
cd $root/eppm_b200/csrc
for f in context prepare patchmatch consistency refine legacy_abi legacy_inplace eval tiled subpix bao_class; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $extra -I../../include -c $f.cu -o $obj/$f.o 2> $obj/$f.ptxas.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xlinker -Bsymbolic -o $root/build/ab/libeppm_b200_$name.so $obj/*.o -ldl 2>/dev/null
echo built build/ab/libeppm_b200_$name.so
