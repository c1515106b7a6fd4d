#!/bin/bash
# ncu --set full of the final y and z passes at C5 (540)
mkdir -p gpurun_out
for k in k_shell_zpass k_shell_ypass; do
  tag=r02ao_${k}540
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o gpurun_out/prof_$tag python scripts/ncu_probe.py C5 > gpurun_out/ncu_$tag.log 2>&1
  ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_$tag.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/prof_$tag.csv > gpurun_out/${tag}.txt
  grep -E "Kernel Name|Grid Size|time_duration|dram__bytes|wavefronts_mem_shared.sum|bank_conflicts_pipe_lsu_mem_shared.sum|registers|warps_active|fp64.avg|long_score.*ratio|short_score.*ratio" gpurun_out/${tag}.txt
done
