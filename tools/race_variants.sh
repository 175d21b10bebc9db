#!/bin/bash
# Builds the experiment variants of the operand-pipeline race investigation (profiles/r02_race_experiments.txt)
# into gpurun_tmp/libagp_<name>.so; run them on a GPU box with tools/race_probe.py (AGP_LIB selects the library).
set -e
cd "$(dirname "$0")/.."
b() { name=$1; shift; tools/build_variant.sh $name "$@" > /dev/null & }
b ship
b simple        -DAGP_X_SIMPLE=1
b simple_rel    -DAGP_X_SIMPLE=1 -DAGP_X_RELFENCE=1
b simple_gf     -DAGP_X_SIMPLE=1 -DAGP_X_GFENCE=1
wait
b simple_both   -DAGP_X_SIMPLE=1 -DAGP_X_RELFENCE=1 -DAGP_X_GFENCE=1
b simple_early  -DAGP_X_SIMPLE=1 -DAGP_X_EARLYREL=1
b simple_dbg    -DAGP_X_SIMPLE=1 -DAGP_X_DEBUGOPS=1
b ship_both     -DAGP_X_RELFENCE=1 -DAGP_X_GFENCE=1
wait
ls -la gpurun_tmp/*.so
