for pos in "0,0.3333,0.5" "0,0.25,0.4" "0,0.2,0.35" "0,0.4,0.6" "0,0.15,0.3" "0.1,0.3333,0.5" "0,0.3333,0.45" "0,0.25,0.5" "0,0.1,0.25"; do
  echo "AGP_POS=$pos"; AGP_POS=$pos python tools/time_lml.py --n 2048 --P 64 --reps 20 --check 0
done
for late in 2 3 4 5 6; do echo "AGP_LATE=$late"; AGP_LATE=$late python tools/time_lml.py --n 2048 --P 64 --reps 20 --check 0; done
