#!/bin/bash
# Deterministic assignment with sorted per-tile candidate lists: bit-exactness tests, then the
# deterministic rows of the assignment sweep with the tile kernel and with the merge kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_production.py -m gpu -x -q -k "bit_exact or deterministic or shadow or meshfield or partition" > gpurun_out/r02y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02y_pytest.log; tail -5 gpurun_out/r02y_pytest.log
SWEEP_ONLY=deterministic timeout 600 python scripts/assign_sweep.py > gpurun_out/r02y_sweep_tile.json 2> gpurun_out/r02y_sweep_tile.err; grep deterministic gpurun_out/r02y_sweep_tile.err
TRV_DET_NO_TILE=1 SWEEP_ONLY=deterministic timeout 600 python scripts/assign_sweep.py > gpurun_out/r02y_sweep_merge.json 2> gpurun_out/r02y_sweep_merge.err; grep deterministic gpurun_out/r02y_sweep_merge.err
