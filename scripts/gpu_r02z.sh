#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_exact or shadow" > gpurun_out/r02z_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02z_pytest.log; tail -3 gpurun_out/r02z_pytest.log
SWEEP_ONLY=deterministic timeout 600 python scripts/assign_sweep.py > gpurun_out/r02z_sweep_tile.json 2> gpurun_out/r02z_sweep_tile.err; grep "tsc\|pcs" gpurun_out/r02z_sweep_tile.err
