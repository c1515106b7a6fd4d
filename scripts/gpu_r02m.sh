#!/bin/bash
# Round 2: Gram product on the FP64 tensor cores (k_gram_dmma): unit + production parity,
# C5 share timings, ncu capture of the C5 launch.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_production.py -m gpu -x -q -k "gram or slab or C2 or C5proxy or live or C1 or C3" > gpurun_out/r02m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02m_pytest.log; tail -8 gpurun_out/r02m_pytest.log
for rw in "0 1" "0 8"; do
  timeout 300 python scripts/c5_share_once.py $rw 4 2>&1 | grep share | tee -a gpurun_out/r02m_shares.txt
done
TRV_PROFILE=1 timeout 300 python - <<'PY' 2>&1 | tail -3 | tee gpurun_out/r02m_c5_phases.txt
import sys, json
sys.path.insert(0, '.')
sys.argv = ['x', '0', '1', '3']
from triumvirate_b200 import core
core.profile_enable(True)
exec(open('scripts/c5_share_once.py').read())
print(json.dumps({k: round(v * 1e3, 2) for k, v in core.profile_report().items()}))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gram_dmma -s 1 -c 1 -f -o gpurun_out/prof_r02m_k_gram_dmma python scripts/c5_share_once.py 0 1 2 > gpurun_out/r02m_ncu.log 2>&1
ncu -i gpurun_out/prof_r02m_k_gram_dmma.ncu-rep --page raw --csv > gpurun_out/prof_r02m_k_gram_dmma.csv 2>/dev/null
python - <<'PY'
import csv,re
rows=list(csv.reader(open('gpurun_out/prof_r02m_k_gram_dmma.csv')))
hdr,units=rows[0],rows[1]
pat=re.compile(r"^(Grid Size|Registers Per Thread|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|dram__bytes_read.sum.per_second|dram__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__throughput.avg.pct|lts__throughput.avg.pct|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_(long|short|wait|math|lg|mio|no_inst|branch|barrier)[a-z_]*_per_issue_active.ratio|l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed)$")
for vals in rows[2:]:
    print('----')
    for h,u,v in zip(hdr,units,vals):
        if pat.search(h): print(f"{h:90s} {v} {u}")
PY
