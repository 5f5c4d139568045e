#!/bin/bash
# tools/build_variant.sh NAME "EXTRA FLAGS": builds ctsm_b200/lib/libctsm_b200_NAME.so (experiment aid; select with CTSM_B200_LIB)
set -e
cd "$(dirname "$0")/.."
name=$1; extra=$2
mkdir -p build/$name
for f in ctsm_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ccbin /usr/bin/g++ $extra -c $f -o build/$name/$(basename $f .cu).o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ctsm_b200/lib/libctsm_b200_$name.so build/$name/*.o -ccbin /usr/bin/g++
echo built ctsm_b200/lib/libctsm_b200_$name.so
