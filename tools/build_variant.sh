#!/bin/bash
# Developer A/B builds: tools/build_variant.sh NAME "-DAGP_GF_E=8 ..."  ->  gpurun_tmp/libagp_NAME.so
# (select at run time with AGP_LIB=gpurun_tmp/libagp_NAME.so)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p gpurun_tmp/$NAME
CS=autogp.jl_b200/csrc
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC $*"
for f in agp_kernels agp_fused agp_api; do $NV -c $CS/$f.cu -o gpurun_tmp/$NAME/$f.o & done
g++ -O2 -std=c++17 -fPIC -c $CS/agp_program.cpp -o gpurun_tmp/$NAME/agp_program.o
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_tmp/libagp_$NAME.so gpurun_tmp/$NAME/*.o -lcudart_static -ldl -lrt -lpthread
rm -rf gpurun_tmp/$NAME
echo built gpurun_tmp/libagp_$NAME.so
