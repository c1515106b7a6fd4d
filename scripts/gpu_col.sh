#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "assign or mesh or survey or golden" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for mode in 0 1; do
  echo "TRV_ASSIGN_COL=$mode"
  TRV_ASSIGN_COL=$mode TRV_PROFILE=1 timeout 900 python scripts/run_configs.py C3 2>&1 | tail -2 | cut -c1-330
done
