#!/bin/bash
# Full GPU suite after the plan-retry fix + shell sub-batch count at C2.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02ai_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ai_pytest_gpu.log; tail -4 gpurun_out/r02ai_pytest_gpu.log
for g in 1 2 3 5; do
  echo "groups $g" >> gpurun_out/r02ai_groups.txt
  TRV_SHELL_GROUPS=$g timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-c5 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['ms_per_step'], j['e2e']['value'], j['gpu_launches'])" >> gpurun_out/r02ai_groups.txt
done
cat gpurun_out/r02ai_groups.txt
