#!/bin/bash
mkdir -p gpurun_out
for lv in 1 2 3 4; do
  TRV_STREAM_LEVELS=$lv BENCH_DEBUG=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lv$lv.log 2> gpurun_out/bench_lv$lv.err
  echo "levels=$lv"; tail -1 gpurun_out/bench_lv$lv.err | cut -c1-120
done
timeout 900 python scripts/c5_probe.py > gpurun_out/c5_probe.txt 2>&1; cat gpurun_out/c5_probe.txt
