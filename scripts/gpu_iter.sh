#!/bin/bash
# Iteration loop: GPU parity tests, bench, one ncu capture of the kernel named in $1.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
BENCH_DEBUG=1 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
if [ -n "$1" ]; then bash scripts/ncu_one.sh $1; fi
