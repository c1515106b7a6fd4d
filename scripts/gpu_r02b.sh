#!/bin/bash
# Round 2, second GPU call: tile-owned assignment -- parity tests, sanitizer, A/B sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "assignment or degenerate or survey_against or streamed or meshfield or twopt_against" > gpurun_out/r02b_pytest_assign.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest_assign.log; tail -15 gpurun_out/r02b_pytest_assign.log
SWEEP_ONLY=throughput timeout 600 python scripts/assign_sweep.py > gpurun_out/r02b_sweep_own.json 2> gpurun_out/r02b_sweep_own.err; cat gpurun_out/r02b_sweep_own.err | tail -20
TRV_ASSIGN_LEGACY=1 SWEEP_ONLY=throughput timeout 600 python scripts/assign_sweep.py > gpurun_out/r02b_sweep_legacy.json 2> gpurun_out/r02b_sweep_legacy.err; cat gpurun_out/r02b_sweep_legacy.err | tail -20
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest_gpu.log; tail -8 gpurun_out/r02b_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 900 gpurun_out/r02b_bench.json
