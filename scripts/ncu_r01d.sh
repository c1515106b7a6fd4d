#!/bin/bash
# Round 1d `ncu --set full` captures: C2 kernels, the 3PCF kernels of C4, the deterministic gather.
mkdir -p gpurun_out
cap() {  # kernel-regex script tag
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${4:-1} -c 1 -f -o gpurun_out/prof_$3 python $2 > gpurun_out/ncu_$3.log 2>&1
  ncu -i gpurun_out/prof_$3.ncu-rep --page raw --csv > gpurun_out/prof_$3.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/prof_$3.csv '^(Kernel Name|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|dram__throughput.avg.pct_of_peak_sustained_elapsed|lts__throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed|launch__registers_per_thread|sm__warps_active.avg.pct_of_peak_sustained_active)$' | cut -c1-150
}
cap k_assign_coop scripts/ncu_probe.py k_assign_coop
cap k_sort_scatter scripts/ncu_probe.py k_sort_scatter
cap k_shot_spectrum scripts/ncu_probe.py k_shot_spectrum
cap k_sjl_apply scripts/ncu_c4.py k_sjl_apply 16
cap k_sjl_prepare scripts/ncu_c4.py k_sjl_prepare 3
cap k_gram_fields_tma scripts/ncu_c4.py k_gram_c4 2
cap k_shot_3pcf_bin scripts/ncu_c4.py k_shot_3pcf_bin 2
cap k_assign_gather_warp scripts/ncu_probe_det.py k_assign_gather_warp 1
rm -f gpurun_out/prof_*.ncu-rep
