#!/bin/bash
# One `ncu --set full` capture per hot kernel (second estimator call of scripts/ncu_probe.py).
mkdir -p gpurun_out
for k in k_assign_coop k_shot_radial_hist k_shot_spectrum k_gram_fields k_gather_sorted k_sort_scatter; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python scripts/ncu_probe.py > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
