# A/B of one against two CTAs per SM for the persistent kernel by batch size (developer tool, GPU box)
for cfg in ${CFGS:-"512 24" "512 32" "512 48" "1024 24" "1024 32" "1024 48" "2048 24" "2048 32" "2048 48" "128 4" "128 64" "256 16"}; do
  set -- $cfg
  for c in 2 1; do
    echo -n "ctas/SM=$c "; AGP_CTAS_PER_SM=$c python tools/time_lml.py --n $1 --P $2 --reps 20 --check 0 | cut -c1-80
  done
done
for cfg in ${GCFGS:-"1024 32" "2048 32" "512 32"}; do
  set -- $cfg
  for c in 2 1; do
    echo -n "ctas/SM=$c "; AGP_CTAS_PER_SM=$c python tools/grad_width_sweep.py $1 $2 4 | tail -1
  done
done
