#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1; do
  TRV_NO_OVERLAP=$mode BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_noov$mode.log 2> gpurun_out/bench_noov$mode.err
  echo "TRV_NO_OVERLAP=$mode"; tail -2 gpurun_out/bench_noov$mode.err | cut -c1-200
done
TRV_NO_OVERLAP=0 timeout 600 python scripts/run_configs.py C5 2>&1 | tail -1
