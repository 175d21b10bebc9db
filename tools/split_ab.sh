# look-ahead partial items of the single-launch schedule from block column 2 instead of 3 (developer A/B, GPU box)
for cfg in "384 64" "512 64" "512 16" "640 64" "1024 16" "1024 64" "768 32"; do
  set -- $cfg
  for sf in 3 2; do
    echo -n "split_from=$sf "; AGP_SPLIT_FROM=$sf python tools/time_lml.py --n $1 --P $2 --reps 30 --check 1 | cut -c9-110
  done
done
