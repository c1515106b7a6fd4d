#!/bin/bash
# Radial shot-noise reduction on k_gram_dmma: tests, production parity, bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_production.py -m gpu -x -q -k "radial or shot or C2 or C1 or C5proxy or live or goldens or partition or slab" > gpurun_out/r02aa_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02aa_pytest.log; tail -4 gpurun_out/r02aa_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02aa_bench.json 2> gpurun_out/r02aa_bench.err
python - <<'PY'
import json
j = json.loads(open('gpurun_out/r02aa_bench.json').read().strip().splitlines()[-1])
print("value", j["value"], "e2e", j["e2e"]["value"], "c5", j["c5"]["ms_per_step"], j["result"]["timed_vs_deterministic_max_rel"], j["gpu_launches"])
PY
timeout 300 python scripts/c5_phase_once.py 0 1 2>&1 | grep phases
