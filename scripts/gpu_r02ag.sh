#!/bin/bash
# ncu --set full captures: k_xpass_fused_async<512> and k_shell_zpass<144> (C2), k_shell_zpass<540> (C5)
mkdir -p gpurun_out
bash scripts/ncu_one.sh k_xpass_fused r02ag_k_xpass_async > /dev/null
python scripts/ncu_summary.py gpurun_out/prof_r02ag_k_xpass_async.csv > gpurun_out/r02ag_ncu_k_xpass_async.txt
bash scripts/ncu_one.sh k_shell_zpass r02ag_k_shell_zpass144 > /dev/null
python scripts/ncu_summary.py gpurun_out/prof_r02ag_k_shell_zpass144.csv > gpurun_out/r02ag_ncu_k_shell_zpass144.txt
k=k_shell_zpass; tag=r02ag_k_shell_zpass540
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o gpurun_out/prof_$tag python scripts/ncu_probe.py C5 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_$tag.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/prof_$tag.csv > gpurun_out/r02ag_ncu_k_shell_zpass540.txt
grep -E "Kernel Name|Grid Size|time_duration|dram__bytes|wavefronts_mem_shared.sum|bank_conflicts_pipe_lsu_mem_shared.sum|registers|warps_active|fp64.avg|issue_active|stalled_(long|short|barrier|mio|math|lg|wait|not_sel).*ratio|occupancy_limit" gpurun_out/r02ag_ncu_k_*.txt
