#!/bin/bash
# Developer build of the persistent kernel for ONE CTA per SM: launch bounds (256, 1), no 128-register cap, optionally with one of
# the POTF2 experiments of csrc/experiments/ in place of the shipped one:  tools/build_solo.sh NAME [experiments/agp_chol_potf2_v3.cu.txt]
#   -> gpurun_tmp/libagp_NAME.so   (run with AGP_LIB=gpurun_tmp/libagp_NAME.so AGP_CTAS_PER_SM=1)
set -e
cd "$(dirname "$0")/.."
NAME=$1; POTF2=$2
B=gpurun_tmp/build_$NAME
rm -rf $B; mkdir -p $B
cp autogp.jl_b200/csrc/*.cu autogp.jl_b200/csrc/*.cuh autogp.jl_b200/csrc/*.h autogp.jl_b200/csrc/*.cpp autogp.jl_b200/csrc/Makefile $B/
[ -n "$POTF2" ] && cp autogp.jl_b200/csrc/$POTF2 $B/agp_chol_potf2.cu
mkdir -p gpurun_tmp/include && cp include/agp_b200.h gpurun_tmp/include/
sed -i 's#../../include/agp_b200.h#../include/agp_b200.h#' $B/*.cu $B/*.cpp $B/*.h $B/Makefile
sed -i 's/__launch_bounds__(FT, 2) agp_chol_kernel/__launch_bounds__(FT, 1) agp_chol_kernel/' $B/agp_chol_kernel.cu
sed -i 's/-maxrregcount=128/-maxrregcount=240/' $B/Makefile
make -s -C $B -j8 OUT=../libagp_$NAME.so > $B/build.log 2>&1 || { tail -30 $B/build.log; exit 1; }
grep -A2 "do_potf2\|agp_chol_kernel" $B/build.log | grep -E "Function properties|registers|spill|Compiling" | head -12
echo built gpurun_tmp/libagp_$NAME.so
