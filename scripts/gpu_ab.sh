#!/bin/bash
# A/B of the two throughput assignment kernels (tile-owned shared memory vs the
# particle-wise cooperative scatter) on the C2 bench and on configs C3/C4.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "assign or mesh" > gpurun_out/pytest_assign.log 2>&1; tail -2 gpurun_out/pytest_assign.log
for mode in 0 1; do
  TRV_ASSIGN_TILE=$mode BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tile$mode.log 2> gpurun_out/bench_tile$mode.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_tile$mode.log").read().strip().splitlines()[-1])
print("tile=$mode ms_per_step", d["ms_per_step"], "assign launch s", d["roofline"]["launch_seconds"], "with sort", d["roofline"]["with_sort"]["launch_seconds"], "e2e", d["e2e"]["value"])
PY
  TRV_ASSIGN_TILE=$mode timeout 900 python scripts/run_configs.py C4 C3 > gpurun_out/configs_tile$mode.json 2> gpurun_out/configs_tile$mode.err; cat gpurun_out/configs_tile$mode.json
done
