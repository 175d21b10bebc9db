# particle-group size of the queue order (AGP_PGROUP): step time, then DRAM traffic of the persistent kernel
for g in 64 32 16 8 4 2; do
  echo "AGP_PGROUP=$g"; AGP_PGROUP=$g python tools/time_lml.py --n 2048 --P 64 --reps 20 --check 0
done
for g in 64 16 8 4; do
  AGP_PGROUP=$g AGP_WAIT_TIMEOUT_MS=600000 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:agp_chol_kernel --launch-skip 2 --launch-count 1 python tools/time_lml.py --n 2048 --P 64 --reps 2 --check 0 2>&1 | grep -E "dram__bytes|hit_rate|gpu__time" | sed "s/^/  pgroup $g: /"
done
