#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for mode in 0 1; do
  TRV_NO_OVERLAP=$mode BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_noov$mode.log 2> gpurun_out/bench_noov$mode.err
  echo "TRV_NO_OVERLAP=$mode"; tail -2 gpurun_out/bench_noov$mode.err | cut -c1-260
done
TRV_NO_OVERLAP=0 timeout 600 python scripts/run_configs.py C5 2>&1 | tail -1
TRV_NO_OVERLAP=1 timeout 600 python scripts/run_configs.py C5 2>&1 | tail -1
