#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/perf_probe.py > gpurun_out/probe.log 2>&1; tail -12 gpurun_out/probe.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_assign_scatter|k_shot_spectrum|k_shot_radial_hist|k_gram|k_sort_scatter|k_shell_spectrum' -s 12 -c 10 -o gpurun_out/prof_r01b python scripts/ncu_probe.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
