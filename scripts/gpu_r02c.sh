#!/bin/bash
# Round 2: tile-owned assignment v2 (two warps per task) -- parity, sweep, ncu capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "assignment or degenerate or survey_against or streamed or meshfield or twopt_against" > gpurun_out/r02c_pytest_assign.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest_assign.log; tail -5 gpurun_out/r02c_pytest_assign.log
SWEEP_ONLY=throughput timeout 600 python scripts/assign_sweep.py > gpurun_out/r02c_sweep_own.json 2> gpurun_out/r02c_sweep_own.err; cat gpurun_out/r02c_sweep_own.err | tail -20
bash scripts/ncu_one.sh k_assign_own r02c_k_assign_own
python scripts/ncu_summary.py gpurun_out/prof_r02c_k_assign_own.csv > gpurun_out/r02c_ncu_k_assign_own.txt; cat gpurun_out/r02c_ncu_k_assign_own.txt | head -80
bash scripts/ncu_one.sh k_own_sort r02c_k_own_sort
python scripts/ncu_summary.py gpurun_out/prof_r02c_k_own_sort.csv > gpurun_out/r02c_ncu_k_own_sort.txt; head -30 gpurun_out/r02c_ncu_k_own_sort.txt
