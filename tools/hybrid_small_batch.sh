# hybrid (int8) against FP64 schedule for small batches (developer tool, GPU box)
for cfg in "2048 8" "2048 16" "2048 32" "1792 16" "1536 16" "1280 16" "1024 16" "4096 8"; do
  set -- $cfg
  for m in 0 1; do
    echo -n "AGP_OZAKI=$m "; AGP_OZAKI=$m python tools/time_lml.py --n $1 --P $2 --reps 10 --check 0 | cut -c9-60
  done
done
