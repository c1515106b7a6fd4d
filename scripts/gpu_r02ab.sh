#!/bin/bash
# Fused x pass of the box mesh phase: tests, production parity, bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_production.py -m gpu -x -q -k "fused or C2 or C1 or C5proxy or goldens or radial" > gpurun_out/r02ab_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ab_pytest.log; tail -15 gpurun_out/r02ab_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02ab_bench.json 2> gpurun_out/r02ab_bench.err
tail -3 gpurun_out/r02ab_bench.err
python - <<'PY'
import json
j = json.loads(open('gpurun_out/r02ab_bench.json').read().strip().splitlines()[-1])
print("value", j["value"], "e2e", j["e2e"]["value"], "c5", j["c5"]["ms_per_step"], j["result"]["timed_vs_deterministic_max_rel"], j["gpu_launches"], j["run"])
print("xpass", j.get("roofline_xpass"))
PY
TRV_NO_FUSED_X=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/r02ab_bench_nofused.json 2>/dev/null
python - <<'PY'
import json
j = json.loads(open('gpurun_out/r02ab_bench_nofused.json').read().strip().splitlines()[-1])
print("nofused value", j["value"], "e2e", j["e2e"]["value"])
PY
