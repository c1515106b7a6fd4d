#!/bin/bash
# Gram (DMMA) tile-length sweep on C5 (41 rows) and C2 (21 rows), after the unit tests.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gram" > gpurun_out/r02o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02o_pytest.log; tail -5 gpurun_out/r02o_pytest.log
for t in 320 256 192; do
  echo "C5 tile $t" | tee -a gpurun_out/r02o_sweep.txt
  TRV_GRAM_TILE=$t timeout 300 python scripts/c5_phase_once.py 0 1 2>&1 | grep phases | tee -a gpurun_out/r02o_sweep.txt
done
for t in 320 256 128; do
  echo "C2 tile $t" | tee -a gpurun_out/r02o_sweep.txt
  TRV_GRAM_TILE=$t timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c5 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', j['ms_per_step'], 'e2e', j['e2e']['value'])" | tee -a gpurun_out/r02o_sweep.txt
done
