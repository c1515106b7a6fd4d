#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
BENCH_DEBUG=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 600 python scripts/perf_probe.py > gpurun_out/probe.log 2>&1; tail -12 gpurun_out/probe.log
timeout 900 python scripts/run_configs.py C1 C4 C3 C5 > gpurun_out/configs.log 2>&1; tail -3 gpurun_out/configs.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc; free -g | head -2
