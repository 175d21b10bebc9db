set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err
AGP_WAIT_TIMEOUT_MS=600000 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_ncu_launches.csv python tools/time_lml.py --n 2048 --P 64 --reps 3 --check 0 > gpurun_out/r02_ncu_launches.log 2>&1
AGP_WAIT_TIMEOUT_MS=600000 ncu --set full --clock-control none --import-source on -k regex:agp_chol_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_chol python tools/time_lml.py --n 2048 --P 64 --reps 2 --check 0 > gpurun_out/r02_ncu_chol.log 2>&1
AGP_WAIT_TIMEOUT_MS=600000 ncu --set full --clock-control none -k regex:agp_gramfill --launch-skip 2 --launch-count 1 -o gpurun_out/r02_gramfill python tools/time_lml.py --n 2048 --P 64 --reps 2 --check 0 > gpurun_out/r02_ncu_gramfill.log 2>&1
ls -la gpurun_out/r02_chol.ncu-rep gpurun_out/r02_gramfill.ncu-rep
tail -2 gpurun_out/r02_ncu_chol.log
