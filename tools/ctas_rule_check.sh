# the by-shape rule (chol_ctas, agp_api.cu) against both forced settings (developer tool, GPU box)
for cfg in "128 64" "256 16" "512 16" "512 32" "512 48" "512 64" "1024 16" "1024 32" "1024 48" "1536 16" "2048 8" "2048 24" "2048 32" "2048 64"; do
  set -- $cfg
  for c in rule 2 1; do
    if [ $c = rule ]; then unset AGP_CTAS_PER_SM; else export AGP_CTAS_PER_SM=$c; fi
    echo -n "ctas/SM=$c "; python tools/time_lml.py --n $1 --P $2 --reps 20 --check 0 | cut -c9-60
  done
done
unset AGP_CTAS_PER_SM
for cfg in "512 8" "512 16" "1024 8" "1024 16" "1024 24" "2048 8" "2048 16"; do
  set -- $cfg
  for c in rule 2 1; do
    if [ $c = rule ]; then unset AGP_CTAS_PER_SM; else export AGP_CTAS_PER_SM=$c; fi
    echo -n "ctas/SM=$c "; python tools/grad_width_sweep.py $1 $2 4 | tail -1 | cut -c1-110
  done
done
