"""Print the headline metrics of a `ncu --page raw --csv` export (wide format)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
  r"^(Kernel Name|Grid Size|Block Size|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum"
  r"|lts__t_sectors_srcunit_tex_op_red.sum|lts__t_sectors.sum|lts__throughput.avg.pct_of_peak_sustained_elapsed"
  r"|l1tex__throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed"
  r"|dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active"
  r"|launch__registers_per_thread|launch__occupancy_limit.*|launch__shared_mem.*|smsp__inst_executed.sum"
  r"|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.*sum|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"
  r"|l1tex__data_pipe_lsu_wavefronts_mem_shared_op_(ld|st).sum|smsp__issue_active.avg.pct.*|sm__inst_executed_pipe_fp64.*pct.*"
  r"|smsp__average_warp.*_per_issue_active.*|smsp__average_warps_issue_stalled_.*_per_issue_active.*|sm__cycles_elapsed.max|smsp__cycles_active.avg"
  r"|smsp__pcsamp_warps_issue_stalled_[a-z_]+)$")
out = []
for h, u, v in zip(hdr, units, vals):
    if pat.search(h):
        out.append((h, v, u))
for h, v, u in out:
    print(f"{h:90s} {v} {u}")
