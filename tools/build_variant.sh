#!/bin/bash
# tools/build_variant.sh NAME "-DGRAD_MIN_BLOCKS=16 ..."  -> openkeonspark_b200/variants/libokb200_NAME.so  (A/B kernel builds; run with OKB200_LIB=...)
set -e
cd "$(dirname "$0")/../openkeonspark_b200"
mkdir -p variants/obj_$1
for f in abi.cpp loader.cpp sampler.cu radix.cu train.cu chunk.cu score.cu tc.cu transr.cu transr_tc.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-fno-fast-math,-ffp-contract=off -x cu -rdc=false $2 -c csrc/$f -o variants/obj_$1/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o variants/libokb200_$1.so variants/obj_$1/*.o -gencode arch=compute_100a,code=sm_100a -lpthread
rm -rf variants/obj_$1
echo variants/libokb200_$1.so
