#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shell_xdft -s 3 -c 2 -f -o gpurun_out/prof_r02i_k_shell_xdft python scripts/c5_share_once.py 0 8 > gpurun_out/r02i_ncu.log 2>&1
ncu -i gpurun_out/prof_r02i_k_shell_xdft.ncu-rep --page raw --csv > gpurun_out/prof_r02i_k_shell_xdft.csv 2>/dev/null
python - <<'PY'
import csv,re
rows=list(csv.reader(open('gpurun_out/prof_r02i_k_shell_xdft.csv')))
hdr,units=rows[0],rows[1]
pat=re.compile(r"^(Grid Size|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|lts__t_sectors.sum|l1tex__throughput.avg.pct|lts__throughput.avg.pct|sm__warps_active.avg.pct|smsp__inst_executed.sum|smsp__issue_active.avg.pct|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|l1tex__t_sector_hit_rate.pct|smsp__average_warps_issue_stalled_(long|short|wait|math|lg|mio|no_inst|branch|barrier)[a-z_]*_per_issue_active.ratio|l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed|sm__inst_executed_pipe_fp64.sum)$")
for vals in rows[2:]:
    print('----')
    for h,u,v in zip(hdr,units,vals):
        if pat.search(h): print(f"{h:90s} {v} {u}")
PY
