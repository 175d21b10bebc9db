# Gram units as queue items against the Gram fill launch, small batches (developer A/B, GPU box)
for cfg in "768 16" "1024 16" "1024 8" "1280 16" "1536 16" "1024 32" "768 32"; do
  set -- $cfg
  for f in 0 1; do
    echo -n "AGP_FUSE_GRAM=$f "; AGP_OZAKI=0 AGP_FUSE_GRAM=$f python tools/time_lml.py --n $1 --P $2 --reps 30 --check 0 | cut -c9-70
  done
done
