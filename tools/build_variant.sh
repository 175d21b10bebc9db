#!/bin/bash
# Developer A/B builds: tools/build_variant.sh NAME "-DAGP_GF_E=8 ..."  ->  gpurun_tmp/libagp_NAME.so
# (select at run time with AGP_LIB=gpurun_tmp/libagp_NAME.so)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
B=gpurun_tmp/build_$NAME
rm -rf $B; mkdir -p $B
cp autogp.jl_b200/csrc/*.cu autogp.jl_b200/csrc/*.cuh autogp.jl_b200/csrc/*.h autogp.jl_b200/csrc/*.cpp autogp.jl_b200/csrc/Makefile $B/
mkdir -p $B/../../include_link && true
# the sources include ../../include/agp_b200.h relative to csrc/: give the copy the same view
mkdir -p gpurun_tmp/include && cp include/agp_b200.h gpurun_tmp/include/ 2>/dev/null || true
sed -i 's#../../include/agp_b200.h#../include/agp_b200.h#' $B/*.cu $B/*.cpp $B/*.h $B/Makefile
make -s -C $B -j8 OUT=../libagp_$NAME.so EXTRA="$*" > $B/build.log 2>&1 || { tail -20 $B/build.log; exit 1; }
rm -rf $B
echo built gpurun_tmp/libagp_$NAME.so
