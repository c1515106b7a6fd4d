#!/bin/bash
# Final state: all GPU tests, smoke, bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02am_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02am_pytest_gpu.log; tail -3 gpurun_out/r02am_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02am_smoke.log 2>&1; tail -2 gpurun_out/r02am_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02am_bench.json 2> gpurun_out/r02am_bench.err
python - <<'PY'
import json
j = json.loads(open('gpurun_out/r02am_bench.json').read().strip().splitlines()[-1])
print("value", j["value"], "e2e", j["e2e"]["value"], "c5", j["c5"]["ms_per_step"], j["result"]["timed_vs_deterministic_max_rel"], j["gpu_launches"], j["result"]["sha256_deterministic_step"][:12])
PY
