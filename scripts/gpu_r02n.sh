#!/bin/bash
# Gram (DMMA) with the producer-warp stage ring: unit tests, then stages x tile sweep on C5.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gram" > gpurun_out/r02n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02n_pytest.log; tail -5 gpurun_out/r02n_pytest.log
for cfg in "128 6" "128 4" "128 3" "128 2" "256 2"; do
  set -- $cfg
  echo "tile $1 stages $2" | tee -a gpurun_out/r02n_sweep.txt
  TRV_GRAM_TILE=$1 TRV_GRAM_STAGES=$2 timeout 300 python scripts/c5_phase_once.py 0 1 2>&1 | grep phases | tee -a gpurun_out/r02n_sweep.txt
done
