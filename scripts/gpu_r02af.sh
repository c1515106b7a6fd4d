#!/bin/bash
# Hand-written pruned z pass: tests, production parity, bench (C2 + C5), C5 phases.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_production.py -m gpu -x -q -k "z_pass or fused or pruned or C2 or C1 or C5proxy or goldens or slab" > gpurun_out/r02af_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02af_pytest.log; tail -15 gpurun_out/r02af_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02af_bench.json 2> gpurun_out/r02af_bench.err
tail -3 gpurun_out/r02af_bench.err
python - <<'PY'
import json
j = json.loads(open('gpurun_out/r02af_bench.json').read().strip().splitlines()[-1])
print("value", j["value"], "e2e", j["e2e"]["value"], "c5", j["c5"]["ms_per_step"], j["result"]["timed_vs_deterministic_max_rel"], j["gpu_launches"], j["run"])
print("xpass", j.get("roofline_xpass"))
PY
timeout 300 python scripts/c5_phase_once.py 0 1 2>&1 | grep phases
TRV_NO_ZPASS=1 timeout 300 python scripts/c5_phase_once.py 0 1 2>&1 | grep phases
