#!/bin/bash
# quick A/B: throughput sweep of the tile-owned assignment + one ncu capture
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "assignment or degenerate" > gpurun_out/r02d_pytest_assign.log 2>&1; tail -3 gpurun_out/r02d_pytest_assign.log
SWEEP_ONLY=throughput timeout 600 python scripts/assign_sweep.py > gpurun_out/r02d_sweep_own.json 2> gpurun_out/r02d_sweep_own.err; grep "uniform\|lognormal/pcs" gpurun_out/r02d_sweep_own.err
bash scripts/ncu_one.sh k_assign_own r02d_k_assign_own > /dev/null
python scripts/ncu_summary.py gpurun_out/prof_r02d_k_assign_own.csv | grep -E "time_duration|wavefronts_mem_shared|l1tex__throughput|warps_active|inst_executed.sum|issue_active|dram__bytes"
