#!/bin/bash
# One GPU-box round: tests, bench, ncu launch list + full captures. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
python bench.py --steps 50 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.json
if [ "$1" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:agp_update -s 24 -c 2 -o gpurun_out/prof_update -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_update.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:agp_potf2 -s 24 -c 1 -o gpurun_out/prof_potf2 -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_potf2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:agp_trsm -s 20 -c 1 -o gpurun_out/prof_trsm -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_trsm.log 2>&1
fi
ls -la gpurun_out
